#!/usr/bin/env python
"""Host-side profile (cProfile) of the zero-edit drop-in path: 8 views x 2 GaussianRasterizer calls + backward."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from dreammesh4d_b200 import rasterizer as R

dev = torch.device("cuda", 0)
scene, graph, node = bench.build_scene(False)
cams = bench.build_cameras(0)
gs = bench.gaussian_sets_gpu(scene, graph, node, dev)
H, W, VIEWS = bench.H, bench.W, bench.VIEWS
V, PV, campos, tanx, tany = cams
Vd, PVd, cd = V.to(dev), PV.to(dev), campos.to(dev)
bg = torch.ones(3, device=dev)
g = torch.Generator().manual_seed(0)
gC, gD, gA = (torch.randn(VIEWS, c, H, W, generator=g).to(dev) for c in (3, 1, 1))
per_view = [[gs["means3D"][v].clone().requires_grad_(True), gs["rotations"][v].clone().requires_grad_(True), gs["normals"][v].clone().requires_grad_(True)] for v in range(VIEWS)]
shared = [gs[k].clone().requires_grad_(True) for k in ("scales", "opacities", "colors")]


def step(normal_grads):
    outs, grads = [], []
    for v in range(VIEWS):
        m, q, nrm = per_view[v]
        s = R.GaussianRasterizationSettings(H, W, float(tanx[v]), float(tany[v]), bg, 1.0, Vd[v], PVd[v], 0, cd[v], False, False)
        rast = R.GaussianRasterizer(s)
        m2d = torch.zeros_like(m, requires_grad=True)
        c, radii, d, a = rast(means3D=m, means2D=m2d, opacities=shared[1], colors_precomp=shared[2], scales=shared[0], rotations=q)
        n, _, _, _ = rast(means3D=m, means2D=torch.zeros_like(m), opacities=shared[1], colors_precomp=nrm, scales=shared[0], rotations=q)
        outs += [c, d, a]
        grads += [gC[v], gD[v], gA[v]]
        if normal_grads:
            outs.append(n)
            grads.append(gC[v])
    torch.autograd.backward(outs, grads)
    for t in shared + [x for pv in per_view for x in pv]:
        t.grad = None


for ng in (False, True):
    for _ in range(3):
        step(ng)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step(ng)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"normal_grads={ng}: host enqueue {1e3 * (t1 - t0) / 10:.2f} ms/step, with GPU drain {1e3 * (t2 - t0) / 10:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step(True)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
