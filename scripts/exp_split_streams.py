#!/usr/bin/env python
"""EXPERIMENT (not product code; first attempt failed on a capture detail, fixed but not re-run — no GPU budget left): overlap the latency-bound kernels of one half of a step's views
(preprocess, scan, scatter, tile sort, preprocess backward: ~0.65 ms of the 2.06 ms step at C3, issue slots ~50 % busy)
with the issue-bound render kernels of the other half, by rasterizing the two halves of the view batch through the
public API on two streams and capturing both into one CUDA graph (a forked graph).  Prints ms/step of the single
8-view call and of the 2 x 4-view split.  Usage on a GPU box:  python scripts/exp_split_streams.py [n_splits]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from dreammesh4d_b200 import rasterizer as R  # noqa: E402


def main():
    n_splits = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda", 0)
    H, W, VIEWS = bench.H, bench.W, bench.VIEWS
    scene, graph, node = bench.build_scene(False)
    V, PV, campos, tanx, tany = bench.build_cameras(0)
    gs = bench.gaussian_sets_gpu(scene, graph, node, dev)
    inp = {k: gs[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "colors")}
    g = torch.Generator().manual_seed(0)
    gC, gD, gA = (torch.randn(VIEWS, c, H, W, generator=g).to(dev) for c in (3, 1, 1))

    def vp_for(views):
        idx = torch.tensor(views)
        return R.make_view_params(V[idx].to(dev), PV[idx].to(dev), campos[idx].to(dev), tanx[idx], tany[idx],
                                  torch.ones(len(views), 3, device=dev), set_index=torch.arange(len(views)))

    st = []
    with torch.no_grad():
        R.rasterize_batch(inp["means3D"], inp["opacities"], inp["scales"], inp["rotations"], inp["colors"],
                          vp_for(list(range(VIEWS))), H, W, distinct_sets=True, state_out=st)
    cap = int(st[0].status()[0] * 1.25) + 4096
    groups = [list(range(VIEWS))[i::n_splits] for i in range(n_splits)]
    vps = [vp_for(gr) for gr in groups]
    vp_all = vp_for(list(range(VIEWS)))
    streams = [torch.cuda.Stream() for _ in groups]
    ixs = [torch.tensor(gr, device=dev) for gr in groups]        # built outside the capture (no pageable H2D inside a graph)

    def one_call():
        c, _, d, a = R.rasterize_batch(inp["means3D"], inp["opacities"], inp["scales"], inp["rotations"], inp["colors"],
                                       vp_all, H, W, capacity=cap, distinct_sets=True)
        torch.autograd.backward([c, d, a], [gC, gD, gA])
        for t in inp.values():
            t.grad = None

    def split_call():
        cur = torch.cuda.current_stream()
        for s, ix, vp in zip(streams, ixs, vps):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                m, r = inp["means3D"].detach()[ix].requires_grad_(True), inp["rotations"].detach()[ix].requires_grad_(True)
                sh = [inp[k].detach().requires_grad_(True) for k in ("opacities", "scales", "colors")]
                c, _, d, a = R.rasterize_batch(m, sh[0], sh[1], r, sh[2], vp, H, W, capacity=cap, distinct_sets=True)
                torch.autograd.backward([c, d, a], [gC[ix], gD[ix], gA[ix]])
        for s in streams:
            cur.wait_stream(s)

    def timed(fn, n=20):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph_):
            fn()
        graph_.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            graph_.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    print(f"single {VIEWS}-view call: {timed(one_call):.3f} ms/step")
    print(f"{n_splits} x {VIEWS // n_splits}-view calls on {n_splits} streams: {timed(split_call):.3f} ms/step")


if __name__ == "__main__":
    main()
