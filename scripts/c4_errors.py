#!/usr/bin/env python
"""Diagnostic: relative L-inf errors of the rasterizer against the CPU oracle at the C4 size (1M Gaussians,
1024x1024), one view, forward images and every gradient.  DM4D_LIB_PATH selects the library variant."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dreammesh4d_b200 import rasterizer as R, synthetic  # noqa: E402
from tests import helpers as Hh  # noqa: E402
from tests.test_raster_parity_gpu import run_oracle  # noqa: E402

P, H, W = 1_000_000, 1024, 1024
means, scales, rots, opac, cols = synthetic.random_gaussians(P, seed=0)
V, PV, campos, tanx, tany = Hh.cameras(16, seed=3)
bg = torch.ones(3)
for v in (0, 11):
    o = run_oracle(P, H, W, means, scales, rots, opac, cols, V[v], PV[v], tanx[v], tany[v], bg)
    ok = torch.from_numpy(~o.ambiguous)[None]
    t = lambda x: x.cuda().requires_grad_(True)
    tm, ts, tr, to_, tc = t(means), t(scales), t(rots), t(opac), t(cols)
    vp = R.make_view_params(V[v:v + 1].cuda(), PV[v:v + 1].cuda(), campos[v:v + 1].cuda(), tanx[v:v + 1], tany[v:v + 1], bg[None].cuda())
    color, radii, depth, alpha = R.rasterize_batch(tm, to_, ts, tr, tc, vp, H, W)
    g = torch.Generator().manual_seed(1)
    gC = torch.randn(3, H, W, generator=g) * ok
    gD = 0.1 * torch.randn(1, H, W, generator=g) * ok
    gA = torch.randn(1, H, W, generator=g) * ok
    ref = o.backward(gC.numpy(), gD.numpy(), gA.numpy())
    ((color[0] * gC.cuda()).sum() + (depth[0] * gD.cuda()).sum() + (alpha[0] * gA.cuda()).sum()).backward()
    okn = ok.numpy()
    img = {n: float(np.abs(a.detach().cpu().numpy() - b)[np.broadcast_to(okn, b.shape)].max() / np.abs(b).max())
           for n, a, b in (("color", color[0], o.color), ("depth", depth[0], o.depth), ("alpha", alpha[0], o.alpha))}
    grads = {n: Hh.rel_linf(tt.grad.cpu().numpy(), ref[n]) for n, tt in
             (("means3D", tm), ("scales", ts), ("rotations", tr), ("opacities", to_), ("colors", tc))}
    from oracle.raster_oracle import RasterOracle
    o64 = RasterOracle(P, H, W, 3, "f64")
    o64.forward(means.numpy(), scales.numpy(), rots.numpy(), opac.numpy(), cols.numpy(), V[v].numpy(), PV[v].numpy(),
                float(tanx[v]), float(tany[v]), np.ones(3, np.float32))
    ok64 = ok.numpy() & ~o64.ambiguous[None]
    m = torch.from_numpy(ok64)
    ref64 = o64.backward((gC * m).numpy(), (gD * m).numpy(), (gA * m).numpy())
    ref32 = o.backward((gC * m).numpy(), (gD * m).numpy(), (gA * m).numpy())
    for tt in (tm, ts, tr, to_, tc):
        tt.grad = None
    ((color[0] * (gC * m).cuda()).sum() + (depth[0] * (gD * m).cuda()).sum() + (alpha[0] * (gA * m).cuda()).sum()).backward()
    g64 = {n: Hh.rel_linf(tt.grad.cpu().numpy(), ref64[n]) for n, tt in
           (("means3D", tm), ("scales", ts), ("rotations", tr), ("opacities", to_), ("colors", tc))}
    o32v64 = {n: Hh.rel_linf(ref32[n], ref64[n]) for n in g64}
    print(f"view {v}: GPU vs f64 oracle", {k: f"{e:.2e}" for k, e in g64.items()}, "| f32 oracle vs f64 oracle", {k: f"{e:.2e}" for k, e in o32v64.items()}, flush=True)
    print(f"view {v}: images", {k: f"{e:.2e}" for k, e in img.items()}, "grads", {k: f"{e:.2e}" for k, e in grads.items()}, flush=True)
