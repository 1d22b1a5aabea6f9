#!/usr/bin/env python
"""Worst relative L-inf gradient error of the rasterizer backward against the CPU oracle at C3 (every view) — the
numbers behind the 1e-3 tolerance of tests/test_full_size_gpu.py.  Test infrastructure (imports the oracle)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from dreammesh4d_b200 import rasterizer as R
from tests import helpers as Hh
from tests.test_raster_parity_gpu import check_view, run_oracle

dev = "cuda"
scene, graph, node = bench.build_scene(False)
V, PV, campos, tanx, tany = bench.build_cameras(0)
gs = bench.gaussian_sets_gpu(scene, graph, node, dev)
n = bench.VIEWS
H = W = 512
t = lambda x: x.detach().clone().requires_grad_(True)
tm, tr, ts, to_, tc = t(gs["means3D"]), t(gs["rotations"]), t(gs["scales"]), t(gs["opacities"]), t(gs["colors"])
vp = R.make_view_params(V.to(dev), PV.to(dev), campos.to(dev), tanx, tany, torch.ones(n, 3), set_index=torch.arange(n))
st = []
color, radii, depth, alpha = R.rasterize_batch(tm, to_, ts, tr, tc, vp, H, W, distinct_sets=True, state_out=st)
g = torch.Generator().manual_seed(7)
gC, gD, gA = torch.randn(n, 3, H, W, generator=g), 0.1 * torch.randn(n, 1, H, W, generator=g), torch.randn(n, 1, H, W, generator=g)
from oracle.raster_oracle import RasterOracle
worst = {"f32": {}, "f64": {}}
amb = []
for v in range(n):
    args = (tm[v].detach().cpu(), ts.detach().cpu(), tr[v].detach().cpu(), to_.detach().cpu(), tc.detach().cpu())
    o = run_oracle(scene.n_gaussians, H, W, *args, V[v], PV[v], tanx[v], tany[v], torch.ones(3))
    o64 = RasterOracle(scene.n_gaussians, H, W, 3, "f64")
    o64.forward(*[a.numpy() for a in args], V[v].numpy(), PV[v].numpy(), float(tanx[v]), float(tany[v]), np.ones(3))
    ok = torch.from_numpy(check_view(o, color[v], radii[v], depth[v], alpha[v], st[0], v) & ~o64.ambiguous & (o64.n_contrib == o.n_contrib))[None]
    amb.append(1.0 - float(ok.float().mean()))
    gC[v] *= ok; gD[v] *= ok; gA[v] *= ok
    refs = {"f32": o.backward(gC[v].numpy(), gD[v].numpy(), gA[v].numpy()), "f64": o64.backward(gC[v].numpy(), gD[v].numpy(), gA[v].numpy())}
    for tt in (tm, tr, ts, to_, tc):
        tt.grad = None
    sel = torch.zeros(n, 1, 1, 1); sel[v] = 1
    ((color * (gC * sel).to(dev)).sum() + (depth * (gD * sel).to(dev)).sum() + (alpha * (gA * sel).to(dev)).sum()).backward(retain_graph=True)
    for name, got in (("means3D", tm.grad[v]), ("rotations", tr.grad[v]), ("scales", ts.grad), ("opacities", to_.grad), ("colors", tc.grad)):
        for prec in refs:
            worst[prec][name] = max(worst[prec].get(name, 0.0), Hh.rel_linf(got.cpu().numpy(), refs[prec][name]))
    worst.setdefault("f32_vs_f64", {})
    for name in ("means3D", "rotations", "scales", "opacities", "colors"):
        worst["f32_vs_f64"][name] = max(worst["f32_vs_f64"].get(name, 0.0), Hh.rel_linf(refs["f32"][name], refs["f64"][name]))
print("C3, 8 views, per view: worst relative L-inf gradient error")
print("  CUDA vs fp32 oracle:", {k: f"{v:.2e}" for k, v in worst["f32"].items()})
print("  CUDA vs fp64 oracle:", {k: f"{v:.2e}" for k, v in worst["f64"].items()})
print("  fp32 oracle vs fp64 oracle:", {k: f"{v:.2e}" for k, v in worst["f32_vs_f64"].items()})
print("excluded-pixel fraction per view (ambiguous in either precision):", [f"{a:.1e}" for a in amb])
