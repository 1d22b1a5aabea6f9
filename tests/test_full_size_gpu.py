"""GPU parity at BASELINE.json's FULL sizes (C3, C4, C5), through the C ABI.

* C3 (100k-face mesh / 300k surface-bound Gaussians / 512x512 / 8 views): the CPU oracle finishes one view in
  well under a second, so every view is compared directly — integer state bit-exact, images <= 1e-4,
  gradients <= 1e-3 (north-star tolerances, tests/helpers.py).
* C4 (1M free Gaussians / 1024x1024 / 16 cameras): two views directly against the oracle, all 16 through
  size-independent properties — per-tile depth sortedness of the instance stream, tile ranges partitioning
  [0, R), transmittance/alpha bounds, linearity of the backward in the image gradients, batched == single-view.
* C5 (200k vertices / 512 control nodes / 600k Gaussians): fused skinning forward + backward against the
  fp64 torch oracle.
"""
import numpy as np
import pytest
import torch

from dreammesh4d_b200 import rasterizer as R
from dreammesh4d_b200 import synthetic
from oracle import skin_oracle as SO
from tests import helpers as Hh
from tests.test_raster_parity_gpu import check_view, run_oracle
from tests.test_skin_parity_gpu import run_both

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_c3_full_size_every_view_vs_oracle():
    n_views, H, W = 8, 512, 512
    scene = synthetic.make_sugar_scene(100_000, g=3)
    graph = synthetic.make_deform_graph(scene.verts, 1000, 4, seed=0)
    node = synthetic.random_node_attrs(n_views, 1000, seed=1)
    with torch.no_grad():
        ref = SO.deform_gaussians(scene, graph, *node)
    P = scene.n_gaussians
    means, rots = ref["means3D"].float().contiguous(), ref["rotations"].float().contiguous()
    scales = torch.cat([torch.full((P, 1), scene.thickness), scene.log_scales.exp()], dim=-1)
    opac, cols = torch.sigmoid(scene.densities), ref["colors"].float().contiguous()
    V, PV, campos, tanx, tany = Hh.cameras(n_views, seed=2)
    bg = torch.ones(3)

    t = lambda x: x.to(DEV).requires_grad_(True)
    tm, ts, tr, to_, tc = t(means), t(scales), t(rots), t(opac), t(cols)
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg[None].expand(n_views, -1).to(DEV),
                            set_index=torch.arange(n_views))
    states = []
    color, radii, depth, alpha = R.rasterize_batch(tm, to_, ts, tr, tc, vp, H, W, distinct_sets=True, state_out=states)
    g = torch.Generator().manual_seed(0)
    gC, gD, gA = torch.randn(n_views, 3, H, W, generator=g), 0.1 * torch.randn(n_views, 1, H, W, generator=g), \
        torch.randn(n_views, 1, H, W, generator=g)
    refs = []
    for v in range(n_views):
        o = run_oracle(P, H, W, means[v], scales, rots[v], opac, cols, V[v], PV[v], tanx[v], tany[v], bg)
        ok = torch.from_numpy(check_view(o, color[v], radii[v], depth[v], alpha[v], states[0], v))[None]
        gC[v] *= ok; gD[v] *= ok; gA[v] *= ok
        refs.append(o.backward(gC[v].numpy(), gD[v].numpy(), gA[v].numpy()))
    assert (alpha > 0.5).float().mean() > 0.2
    ((color * gC.to(DEV)).sum() + (depth * gD.to(DEV)).sum() + (alpha * gA.to(DEV)).sum()).backward()
    for v in range(n_views):      # per-timestamp attribute sets: one gradient slab per view
        assert Hh.rel_linf(tm.grad[v].cpu().numpy(), refs[v]["means3D"]) <= Hh.TOL_GRAD, f"means3D view {v}"
        assert Hh.rel_linf(tr.grad[v].cpu().numpy(), refs[v]["rotations"]) <= Hh.TOL_GRAD, f"rotations view {v}"
    for name, tt in (("colors", tc), ("opacities", to_), ("scales", ts)):   # shared attributes: sum over views
        want = sum(r[name].astype(np.float64) for r in refs)
        assert Hh.rel_linf(tt.grad.cpu().numpy(), want) <= Hh.TOL_GRAD, name


def _view_depths(means, V):
    """View-space z of every Gaussian (row-vector convention of threestudio/utils/ops.py:398-413)."""
    return means @ V[:3, 2] + V[3, 2]


def test_c4_microbench_size_oracle_views_and_properties():
    P, H, W, B = 1_000_000, 1024, 1024, 16
    means, scales, rots, opac, cols = synthetic.random_gaussians(P, seed=0)
    V, PV, campos, tanx, tany = Hh.cameras(B, seed=3)
    bg = torch.ones(3)
    t = lambda x: x.to(DEV).requires_grad_(True)
    tm, ts, tr, to_, tc = t(means), t(scales), t(rots), t(opac), t(cols)
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg[None].expand(B, -1).to(DEV))
    states = []
    color, radii, depth, alpha = R.rasterize_batch(tm, to_, ts, tr, tc, vp, H, W, state_out=states)
    R_total, overflow = states[0].status()
    assert not overflow and R_total > P

    # --- properties over all 16 views ---------------------------------------------------------------
    assert float(alpha.min()) >= 0.0 and float(alpha.max()) <= 1.0
    assert float(color.min()) >= -1e-5 and float(color.max()) <= 1.0 + 1e-5      # colours, bg in [0,1]
    assert bool(torch.isfinite(depth).all())
    r_sum = 0
    for v in range(B):
        ranges, pl, nc = states[0].export_view(v)
        ranges = ranges.long()
        ne = ranges[ranges[:, 1] > ranges[:, 0]]
        # non-empty tile ranges tile [0, R_v) in tile order without gaps or overlaps
        assert int(ne[0, 0]) == 0 and bool((ne[1:, 0] == ne[:-1, 1]).all()) and int(ne[-1, 1]) == pl.numel()
        r_sum += pl.numel()
        # view-space depth ascends inside every tile (fp64 recomputation; slack = a few fp32 ulps of the key)
        z = _view_depths(tm.detach().double(), V[v].double().to(DEV))[pl.long()]
        dz = z[1:] - z[:-1]
        inner = torch.ones(pl.numel() - 1, dtype=torch.bool, device=DEV)
        inner[(ne[:-1, 1] - 1).clamp_min(0)] = False                          # boundaries between tiles
        assert bool((dz >= -4e-6)[inner].all()), f"view {v}: tile not depth-sorted"
        # every Gaussian appears tiles_touched times
        rect_tiles = torch.bincount(pl.long(), minlength=P)
        assert bool(((rect_tiles > 0) == (radii[v] > 0)).all())
        assert int(nc.max()) <= int((ranges[:, 1] - ranges[:, 0]).max())
    assert r_sum == R_total

    # --- two views directly against the oracle -------------------------------------------------------
    gC = torch.zeros(B, 3, H, W); gD = torch.zeros(B, 1, H, W); gA = torch.zeros(B, 1, H, W)
    g = torch.Generator().manual_seed(1)
    refs = {}
    for v in (0, 11):
        o = run_oracle(P, H, W, means, scales, rots, opac, cols, V[v], PV[v], tanx[v], tany[v], bg)
        # ~100 blended Gaussians per pixel: more pixels sit next to a hard threshold than in the surface scenes
        ok = torch.from_numpy(check_view(o, color[v], radii[v], depth[v], alpha[v], states[0], v, max_ambig=6e-3))[None]
        gC[v] = torch.randn(3, H, W, generator=g) * ok
        gD[v] = 0.1 * torch.randn(1, H, W, generator=g) * ok
        gA[v] = torch.randn(1, H, W, generator=g) * ok
        refs[v] = o.backward(gC[v].numpy(), gD[v].numpy(), gA[v].numpy())
    loss = lambda c, d, a, s=1.0: ((c * (s * gC).to(DEV)).sum() + (d * (s * gD).to(DEV)).sum() + (a * (s * gA).to(DEV)).sum())
    leaves = (tm, ts, tr, to_, tc)
    g1 = torch.autograd.grad(loss(color, depth, alpha), leaves, retain_graph=True)
    for name, got in zip(("means3D", "scales", "rotations", "opacities", "colors"), g1):
        want = refs[0][name].astype(np.float64) + refs[11][name]
        assert Hh.rel_linf(got.cpu().numpy(), want) <= Hh.TOL_GRAD, name

    # --- linearity of the backward in the image gradients --------------------------------------------
    h = torch.Generator().manual_seed(2)
    hC = torch.randn(B, 3, H, W, generator=h).to(DEV)
    g2 = torch.autograd.grad((color * hC).sum(), leaves, retain_graph=True)
    g3 = torch.autograd.grad(loss(color, depth, alpha, 2.0) - 3.0 * (color * hC).sum(), leaves)
    for a, b, c in zip(g1, g2, g3):
        want = 2.0 * a.double() - 3.0 * b.double()
        assert float((c.double() - want).abs().max() / want.abs().max().clamp_min(1e-30)) <= 2e-4

    # --- a view rendered alone equals its slice of the batch, bit for bit ----------------------------
    c1, r1, d1, a1 = R.rasterize_batch(tm.detach(), to_.detach(), ts.detach(), tr.detach(), tc.detach(), vp[5:6], H, W)
    assert torch.equal(c1[0], color[5]) and torch.equal(r1[0], radii[5]) and torch.equal(d1[0], depth[5]) and \
        torch.equal(a1[0], alpha[5])


def test_c5_skinning_microbench_size():
    """200k vertices, 512 control nodes (K = 4), a 200k-face subset x 3 Gaussians = 600k."""
    full = synthetic.make_sugar_scene(400_000, g=3)
    assert abs(full.verts.shape[0] - 200_000) < 2_000
    sel = torch.randperm(400_000, generator=torch.Generator().manual_seed(0))[:200_000].sort()[0]
    gsel = (sel[:, None] * 3 + torch.arange(3)[None]).reshape(-1)
    scene = synthetic.SugarScene(full.verts, full.faces[sel].contiguous(), full.bary, full.log_scales[gsel].contiguous(),
                                 full.complex_rot[gsel].contiguous(), full.densities[gsel].contiguous(),
                                 full.sh_dc[gsel].contiguous(), full.thickness, 3)
    assert scene.n_gaussians == 600_000
    graph = synthetic.make_deform_graph(scene.verts, 512, 4)
    node = synthetic.random_node_attrs(2, 512, seed=7)
    run_both(scene, graph, node, "hybrid", seed=3)
