"""GPU parity: libdm4d.so rasterizer (through the C ABI / drop-in module) vs the CPU oracle.

Tolerances are the north-star's: 1e-4 relative L-inf on colour/depth/alpha, 1e-3 on gradients,
bit-exact integer state (radii, tile ranges, sorted instance ids, n_contrib).  Pixels whose
compositing decisions sit within the oracle's ambiguity margin of a hard threshold
(alpha < 1/255, T < 1e-4, power > 0) are excluded and counted.
"""
import math

import numpy as np
import pytest
import torch

from dreammesh4d_b200 import rasterizer as R
from dreammesh4d_b200 import synthetic
from oracle.raster_oracle import RasterOracle
from tests import helpers as Hh

pytestmark = pytest.mark.gpu
DEV = "cuda"
MAX_AMBIG_FRACTION = 2e-3


def run_oracle(P, H, W, means, scales, rots, opac, cols, V, PV, tanx, tany, bg, C=3):
    o = RasterOracle(P, H, W, C, "f32")
    o.forward(means.numpy(), scales.numpy(), rots.numpy(), opac.numpy(), cols.numpy(), V.numpy(), PV.numpy(),
              float(tanx), float(tany), np.asarray(bg, dtype=np.float32))
    return o


def check_view(o: RasterOracle, color, radii, depth, alpha, state, view, check_state=True,
               max_ambig=MAX_AMBIG_FRACTION):
    amb = o.ambiguous
    ok = ~amb
    assert amb.mean() <= max_ambig, f"too many ambiguous pixels: {amb.mean()}"
    np.testing.assert_array_equal(radii.cpu().numpy(), o.radii)
    if check_state:
        ranges, pl, nc = state.export_view(view)
        np.testing.assert_array_equal(ranges.cpu().numpy().astype(np.uint32), o.ranges)
        ids, _ = o.point_list()
        np.testing.assert_array_equal(pl.cpu().numpy().astype(np.uint32), ids)
        np.testing.assert_array_equal(nc.cpu().numpy().astype(np.uint32)[ok], o.n_contrib[ok])
    for name, got, ref in (("color", color, o.color), ("depth", depth, o.depth), ("alpha", alpha, o.alpha)):
        got = got.detach().cpu().numpy()
        m = np.broadcast_to(ok, ref.shape)
        err = np.abs(got - ref)[m].max() / max(np.abs(ref).max(), 1e-30)
        assert err <= Hh.TOL_IMAGE, f"{name}: rel Linf {err}"
    return ok


@pytest.mark.parametrize("P,H,W,seed", [(3000, 100, 130, 0), (20000, 256, 256, 1), (64, 48, 64, 2)])
def test_dropin_single_view_forward_backward(P, H, W, seed):
    means, scales, rots, opac, cols = Hh.random_scene(P, seed)
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=seed + 10)
    bg = torch.tensor([1.0, 1.0, 1.0])
    o = run_oracle(P, H, W, means, scales, rots, opac, cols, V[0], PV[0], tanx[0], tany[0], bg)

    t = lambda x: x.to(DEV).requires_grad_(True)
    tm, ts, tr, to_, tc = t(means), t(scales), t(rots), t(opac), t(cols)
    m2d = torch.zeros(P, 3, device=DEV, requires_grad=True)
    settings = R.GaussianRasterizationSettings(H, W, float(tanx[0]), float(tany[0]), bg.to(DEV), 1.0, V[0].to(DEV),
                                               PV[0].to(DEV), 0, campos[0].to(DEV), False, False)
    states = []
    vp = R.make_view_params(V[:1].to(DEV), PV[:1].to(DEV), campos[:1].to(DEV), tanx[:1], tany[:1], bg[None].to(DEV))
    color, radii, depth, alpha = R.rasterize_batch(tm, to_, ts, tr, tc, vp, H, W, means2D=m2d[None], state_out=states)
    ok = check_view(o, color[0], radii[0], depth[0], alpha[0], states[0], 0)

    # drop-in module gives the same numbers
    c2, r2, d2, a2 = R.GaussianRasterizer(settings)(means3D=tm, means2D=m2d, opacities=to_, colors_precomp=tc,
                                                     scales=ts, rotations=tr)
    assert torch.equal(c2, color[0]) and torch.equal(r2, radii[0]) and torch.equal(d2, depth[0]) and torch.equal(a2, alpha[0])

    g = torch.Generator().manual_seed(seed)
    gC = torch.randn(3, H, W, generator=g) * torch.from_numpy(ok)[None]
    gD = torch.randn(1, H, W, generator=g) * torch.from_numpy(ok)[None]
    gA = torch.randn(1, H, W, generator=g) * torch.from_numpy(ok)[None]
    (color[0] * gC.to(DEV)).sum().add((depth[0] * gD.to(DEV)).sum()).add((alpha[0] * gA.to(DEV)).sum()).backward()
    ref = o.backward(gC.numpy(), gD.numpy(), gA.numpy())
    for name, tt in (("means3D", tm), ("means2D", m2d), ("colors", tc), ("opacities", to_), ("scales", ts),
                     ("rotations", tr)):
        err = Hh.rel_linf(tt.grad.cpu().numpy(), ref[name])
        assert err <= Hh.TOL_GRAD, f"grad {name}: rel Linf {err}"


def test_sugar_sphere_c1_static():
    """BASELINE config 1 geometry: 10k-face sphere, 30k bound Gaussians, 256x256, 1 view."""
    scene = synthetic.make_sugar_scene(10_000, g=3)
    means, scales, rots, opac, cols, normals = Hh.sugar_gaussians(scene)
    P, H, W = means.shape[0], 256, 256
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=5)
    bg = torch.tensor([1.0, 1.0, 1.0])
    o = run_oracle(P, H, W, means, scales, rots, opac, cols, V[0], PV[0], tanx[0], tany[0], bg)
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg[None].to(DEV))
    states = []
    t = lambda x: x.to(DEV).requires_grad_(True)
    tm, ts, tr, to_, tc = t(means), t(scales), t(rots), t(opac), t(cols)
    color, radii, depth, alpha = R.rasterize_batch(tm, to_, ts, tr, tc, vp, H, W, state_out=states)
    ok = check_view(o, color[0], radii[0], depth[0], alpha[0], states[0], 0)
    assert (alpha > 0.5).float().mean() > 0.2     # the sphere is actually in view
    gC = torch.randn(3, H, W) * torch.from_numpy(ok)[None]
    (color[0] * gC.to(DEV)).sum().backward()
    ref = o.backward(gC.numpy())
    for name, tt in (("means3D", tm), ("colors", tc), ("opacities", to_), ("scales", ts), ("rotations", tr)):
        err = Hh.rel_linf(tt.grad.cpu().numpy(), ref[name])
        assert err <= Hh.TOL_GRAD, f"grad {name}: rel Linf {err}"


def test_batched_views_sets_and_fused_six_channels():
    """4 views over 2 attribute sets; the 6-channel fused pass equals two 3-channel passes and the oracle."""
    P, H, W, B, S = 5000, 128, 128, 4, 2
    means, scales, rots, opac, cols = Hh.random_scene(P, 3)
    means2 = means + 0.05 * torch.randn(P, 3, generator=torch.Generator().manual_seed(9))
    nrm = torch.nn.functional.normalize(torch.randn(S, P, 3, generator=torch.Generator().manual_seed(4)), dim=-1)
    V, PV, campos, tanx, tany = Hh.cameras(B, seed=7)
    bg = torch.tensor([1.0, 1.0, 1.0, 1.0, 1.0, 1.0])
    set_idx = torch.tensor([0, 1, 1, 0])
    vp6 = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg[None].expand(B, -1).to(DEV),
                             set_index=set_idx)
    M = torch.stack([means, means2]).to(DEV).requires_grad_(True)
    N = nrm.to(DEV).requires_grad_(True)
    t = lambda x: x.to(DEV).requires_grad_(True)
    ts, tr, to_, tc = t(scales), t(rots), t(opac), t(cols)
    states = []
    col6, radii, depth, alpha = R.rasterize_batch(M, to_, ts, tr, tc, vp6, H, W, colors2=N, state_out=states)
    gC = torch.randn(B, 6, H, W, generator=torch.Generator().manual_seed(5))
    oks, refs = [], []
    for v in range(B):
        s = int(set_idx[v])
        m = [means, means2][s]
        feat = torch.cat([cols, nrm[s]], dim=1)
        o = run_oracle(P, H, W, m, scales, rots, opac, feat, V[v], PV[v], tanx[v], tany[v], bg, C=6)
        ok = check_view(o, col6[v], radii[v], depth[v], alpha[v], states[0], v)
        gC[v] *= torch.from_numpy(ok)[None]
        refs.append(o.backward(gC[v].numpy()))
        oks.append(ok)
    (col6 * gC.to(DEV)).sum().backward()
    # expected: per-set sums over the views that used the set; shared attributes sum over all views
    exp_means = np.zeros((S, P, 3)); exp_n = np.zeros((S, P, 3))
    exp = {k: 0.0 for k in ("scales", "rotations", "opacities")}
    exp_cols = 0.0
    for v in range(B):
        s = int(set_idx[v])
        exp_means[s] += refs[v]["means3D"]
        exp_n[s] += refs[v]["colors"][:, 3:]
        exp_cols = exp_cols + refs[v]["colors"][:, :3]
        for k in exp:
            exp[k] = exp[k] + refs[v][k]
    assert Hh.rel_linf(M.grad.cpu().numpy(), exp_means) <= Hh.TOL_GRAD
    assert Hh.rel_linf(N.grad.cpu().numpy(), exp_n) <= Hh.TOL_GRAD
    assert Hh.rel_linf(tc.grad.cpu().numpy(), exp_cols) <= Hh.TOL_GRAD
    assert Hh.rel_linf(ts.grad.cpu().numpy(), exp["scales"]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(tr.grad.cpu().numpy(), exp["rotations"]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(to_.grad.cpu().numpy(), exp["opacities"]) <= Hh.TOL_GRAD

    # two 3-channel passes (the reference's call pattern) give bit-identical images
    vp3 = vp6.clone()
    c_rgb, _, d3, a3 = R.rasterize_batch(M.detach(), to_.detach(), ts.detach(), tr.detach(), tc.detach(), vp3, H, W)
    c_nrm, _, _, _ = R.rasterize_batch(M.detach(), to_.detach(), ts.detach(), tr.detach(), N.detach(), vp3, H, W)
    assert torch.equal(c_rgb, col6[:, :3]) and torch.equal(c_nrm, col6[:, 3:])
    assert torch.equal(d3, depth) and torch.equal(a3, alpha)


def test_capacity_mode_and_overflow_flag():
    P, H, W = 4000, 96, 96
    means, scales, rots, opac, cols = Hh.random_scene(P, 11)
    V, PV, campos, tanx, tany = Hh.cameras(2, seed=3)
    bg = torch.ones(2, 3)
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg.to(DEV))
    args = [x.to(DEV) for x in (means, opac, scales, rots, cols)]
    st = []
    ref = R.rasterize_batch(*args, vp, H, W, state_out=st)
    n, over = st[0].status()
    assert n > 0 and not over
    st2 = []
    got = R.rasterize_batch(*args, vp, H, W, capacity=n + 1000, state_out=st2)
    assert st2[0].status() == (n, False)
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    st3 = []
    R.rasterize_batch(*args, vp, H, W, capacity=max(n // 2, 1), state_out=st3)
    assert st3[0].status() == (n, True)


def test_empty_and_fully_culled_inputs():
    H, W = 64, 64
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=1)
    bg = torch.tensor([[0.25, 0.5, 0.75]])
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg.to(DEV))
    # every Gaussian behind the camera
    P = 100
    means, scales, rots, opac, cols = Hh.random_scene(P, 1)
    means = means + campos[0] * 2.0
    color, radii, depth, alpha = R.rasterize_batch(*[x.to(DEV) for x in (means, opac, scales, rots, cols)], vp, H, W)
    assert int(radii.abs().sum()) == 0 and float(alpha.abs().max()) == 0.0
    assert torch.allclose(color[0], bg.to(DEV)[0][:, None, None].expand(3, H, W))
    # P = 0
    z = lambda k: torch.zeros(0, k, device=DEV)
    color, radii, depth, alpha = R.rasterize_batch(z(3), z(1), z(3), z(4), z(3), vp, H, W)
    assert radii.shape == (1, 0) and torch.allclose(color[0], bg.to(DEV)[0][:, None, None].expand(3, H, W))


def test_host_streamed_step_matches_direct_call():
    """HostStreamedRasterStep (pinned host buffers, 3-stream software pipeline) returns the same images and
    gradients as a direct device-resident rasterize_batch call, for several pipelined steps."""
    from dreammesh4d_b200.streaming import HostStreamedRasterStep
    P, H, W, B = 3000, 96, 96, 3
    means, scales, rots, opac, cols = Hh.random_scene(P, 21)
    gen = torch.Generator().manual_seed(3)
    M = torch.stack([means + 0.02 * torch.randn(P, 3, generator=gen) for _ in range(B)])
    Rr = torch.nn.functional.normalize(rots[None] + 0.05 * torch.randn(B, P, 4, generator=gen), dim=-1)
    V, PV, campos, tanx, tany = Hh.cameras(B, seed=9)
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, torch.ones(B, 3, device=DEV),
                            set_index=torch.arange(B))
    gC, gD, gA = (torch.randn(B, k, H, W, generator=gen) for k in (3, 1, 1))
    t = lambda x: x.to(DEV).requires_grad_(True)
    dm, dr, ds, do, dc = t(M), t(Rr), t(scales), t(opac), t(cols)
    st = []
    color, radii, depth, alpha = R.rasterize_batch(dm, do, ds, dr, dc, vp, H, W, distinct_sets=True, state_out=st)
    torch.autograd.backward([color, depth, alpha], [gC.to(DEV), gD.to(DEV), gA.to(DEV)])
    n, _ = st[0].status()
    pin = lambda x: x.contiguous().pin_memory()
    host = {"means": pin(M), "rots": pin(Rr), "scales": pin(scales), "opac": pin(opac), "cols": pin(cols)}
    step = HostStreamedRasterStep(host, {"gC": pin(gC), "gD": pin(gD), "gA": pin(gA)}, vp, H, W, capacity=n + 1024)
    for _ in range(4):
        step.step()
    step.drain()
    torch.cuda.synchronize()
    assert not step.overflowed()
    assert torch.equal(step.out["color"], color.detach().cpu()) and torch.equal(step.out["alpha"], alpha.detach().cpu())
    for name, ref in (("means", dm), ("rots", dr), ("scales", ds), ("opac", do), ("cols", dc)):
        assert Hh.rel_linf(step.out[name].numpy(), ref.grad.cpu().numpy()) <= 1e-5, name


def test_dropin_shs_path_equals_precomputed_colours():
    """``shs=`` (the reference's predict_step, degree 0; here also degree 2): colours evaluated by sh_to_rgb on the device
    and rasterized — identical to passing the same colours as colors_precomp, and gradients reach the coefficients."""
    P, H, W = 2000, 64, 80
    means, scales, rots, opac, cols = Hh.random_scene(P, 5)
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=3)
    bg = torch.ones(3)
    for deg in (0, 2):
        g = torch.Generator().manual_seed(deg)
        shs = (torch.randn(P, (deg + 1) ** 2, 3, generator=g) * 0.5).to(DEV).requires_grad_(True)
        settings = R.GaussianRasterizationSettings(H, W, float(tanx[0]), float(tany[0]), bg.to(DEV), 1.0, V[0].to(DEV),
                                                   PV[0].to(DEV), deg, campos[0].to(DEV), False, False)
        rast = R.GaussianRasterizer(settings)
        d = lambda x: x.to(DEV)
        c1, r1, d1, a1 = rast(means3D=d(means), means2D=torch.zeros(P, 3, device=DEV), opacities=d(opac), shs=shs,
                              scales=d(scales), rotations=d(rots))
        pre = R.sh_to_rgb(shs.detach(), d(means), campos[0].to(DEV), deg)
        c2, r2, d2, a2 = rast(means3D=d(means), means2D=torch.zeros(P, 3, device=DEV), opacities=d(opac), colors_precomp=pre,
                              scales=d(scales), rotations=d(rots))
        assert torch.equal(c1, c2) and torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(a1, a2)
        c1.sum().backward()
        assert shs.grad is not None and float(shs.grad.abs().max()) > 0


@pytest.mark.parametrize("P", [511, 512, 513, 1023, 1024, 1025, 2048, 2049, 4096, 4097, 8192, 8193, 16384, 16385, 21000])
def test_tile_sort_tier_boundaries(P):
    """One 16x16 image = one tile holding every Gaussian: segment lengths on both sides of every tier of the per-tile
    sort (128x8, 512x8, 1024x8, 1024x16 keys in shared memory, chunked merging through global memory above 16384).
    Sorted instance ids and tile ranges bit-exact against the oracle; duplicated depths exercise the (depth, id) order."""
    g = torch.Generator().manual_seed(P)
    means = 0.15 * (torch.rand(P, 3, generator=g) - 0.5)
    means[: P // 7, 2] = means[P // 7: 2 * (P // 7), 2][: P // 7]          # ties in view-space depth for an axis-aligned camera
    scales = torch.exp(torch.rand(P, 3, generator=g) * (math.log(0.03) - math.log(0.005)) + math.log(0.005))
    rots = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1)
    opac = torch.rand(P, 1, generator=g) * 0.5 + 0.05
    cols = torch.rand(P, 3, generator=g)
    H = W = 16
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=3)
    bg = torch.tensor([0.0, 0.0, 0.0])
    o = run_oracle(P, H, W, means, scales, rots, opac, cols, V[0], PV[0], tanx[0], tany[0], bg)
    vp = R.make_view_params(V.to(DEV), PV.to(DEV), campos.to(DEV), tanx, tany, bg[None].to(DEV))
    states = []
    d = lambda x: x.to(DEV)
    color, radii, depth, alpha = R.rasterize_batch(d(means), d(opac), d(scales), d(rots), d(cols), vp, H, W, state_out=states)
    assert int((radii[0] > 0).sum()) == P                                  # nothing culled: the tile's segment has P keys
    check_view(o, color[0], radii[0], depth[0], alpha[0], states[0], 0, max_ambig=0.05)


def test_cov3d_precomp_matches_scale_rotation_path():
    """``cov3D_precomp=`` of the drop-in module: with Sigma = L L^T (L = R(q) diag(s)) evaluated in the kernel's own
    operation order (separately rounded fp32), the integer state and the images are IDENTICAL to the scales / rotations
    call (and hence to the oracle), and dL/dcov3D pulled back through Sigma(s, q) reproduces dL/dscales, dL/drotations."""
    P, H, W = 4000, 96, 128
    means, scales, rots, opac, cols = Hh.random_scene(P, 21)
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=31)
    bg = torch.tensor([0.3, 0.2, 0.1])

    def sigma(s, q):                      # the kernel's formula and order (raster_project.cuh: gaussian_sigma)
        r, x, y, z = q.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
        L = R * s[:, None, :]
        S = lambda a, b: L[:, a, 0] * L[:, b, 0] + L[:, a, 1] * L[:, b, 1] + L[:, a, 2] * L[:, b, 2]
        return torch.stack([S(0, 0), S(0, 1), S(0, 2), S(1, 1), S(1, 2), S(2, 2)], dim=-1)

    d = lambda x: x.to(DEV)
    settings = R.GaussianRasterizationSettings(H, W, float(tanx[0]), float(tanx[0] * 0 + tany[0]), d(bg), 1.0, d(V[0]), d(PV[0]), 0,
                                               d(campos[0]), False, False)
    ts, tq = d(scales).requires_grad_(True), d(rots).requires_grad_(True)
    tm, to_, tc = d(means).requires_grad_(True), d(opac), d(cols)
    c1, r1, d1, a1 = R.GaussianRasterizer(settings)(means3D=tm, means2D=None, opacities=to_, colors_precomp=tc, scales=ts, rotations=tq)
    cov = sigma(scales, rots)             # CPU fp32, op by op like the kernel (-fmad=false)
    tcov = d(cov).requires_grad_(True)
    tm2 = d(means).requires_grad_(True)
    c2, r2, d2, a2 = R.GaussianRasterizer(settings)(means3D=tm2, means2D=None, opacities=to_, colors_precomp=tc, cov3D_precomp=tcov)
    assert torch.equal(r1, r2)
    for x, y in ((c1, c2), (d1, d2), (a1, a2)):
        assert Hh.rel_linf(y.detach().cpu().numpy(), x.detach().cpu().numpy()) <= 1e-6
    g = torch.Generator().manual_seed(5)
    gC, gD, gA = d(torch.randn(3, H, W, generator=g)), d(torch.randn(1, H, W, generator=g)), d(torch.randn(1, H, W, generator=g))
    ((c1 * gC).sum() + (d1 * gD).sum() + (a1 * gA).sum()).backward()
    ((c2 * gC).sum() + (d2 * gD).sum() + (a2 * gA).sum()).backward()
    assert Hh.rel_linf(tm2.grad.cpu().numpy(), tm.grad.cpu().numpy()) <= 1e-5
    # pull dL/dcov3D back through Sigma(s, q) in fp64
    s64, q64 = scales.double().requires_grad_(True), rots.double().requires_grad_(True)
    (sigma(s64, q64) * tcov.grad.cpu().double()).sum().backward()
    assert Hh.rel_linf(s64.grad.numpy(), ts.grad.cpu().double().numpy()) <= Hh.TOL_GRAD
    assert Hh.rel_linf(q64.grad.numpy(), tq.grad.cpu().double().numpy()) <= Hh.TOL_GRAD
