"""Property tests of the rasterizer: invariants that hold for ANY scene, checked with hypothesis on
random small scenes — on the CPU oracle (always) and on the CUDA path through the public API (``-m gpu``).

  * partition of unity: composited weights + final transmittance = 1, so a scene of ONE colour over a background of the
    same colour renders exactly that colour (up to fp32 rounding), whatever the geometry;
  * permutation invariance: the input order of the Gaussians does not matter (the sort key is (tile, depth, id) and
    depths are distinct with probability 1);
  * monotonicity: growing every Gaussian never shrinks a radius, a tile rectangle or a pixel's alpha lower bound 0;
  * zero opacity renders the background and sends no gradient anywhere;
  * linearity of the backward in the image gradients.
"""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle.raster_oracle import RasterOracle
from tests import helpers as Hh

H, W = 40, 56
COMMON = dict(deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def scene(P, seed, lo=0.01, hi=0.08):
    means, scales, rots, opac, cols = Hh.random_scene(P, seed=seed, scale_lo=lo, scale_hi=hi)
    V, PV, campos, tanx, tany = Hh.cameras(1, seed=seed + 1)
    return (means, scales, rots, opac, cols), (V[0], PV[0], float(tanx[0]), float(tany[0]))


def oracle_render(s, cam, bg, cols=None):
    means, scales, rots, opac, c = s
    o = RasterOracle(means.shape[0], H, W, 3, "f32")
    o.forward(means.numpy(), scales.numpy(), rots.numpy(), opac.numpy(), (c if cols is None else cols).numpy(), cam[0].numpy(),
              cam[1].numpy(), cam[2], cam[3], np.asarray(bg, np.float32))
    return o


@settings(max_examples=12, **COMMON)
@given(P=st.integers(1, 400), seed=st.integers(0, 10_000), r=st.floats(0, 1), g=st.floats(0, 1), b=st.floats(0, 1))
def test_oracle_partition_of_unity(P, seed, r, g, b):
    s, cam = scene(P, seed)
    col = torch.tensor([r, g, b], dtype=torch.float32)
    o = oracle_render(s, cam, col.numpy(), cols=col.expand(P, 3).contiguous())
    assert np.abs(o.color - col.numpy()[:, None, None]).max() <= 2e-6
    assert o.alpha.min() >= 0 and o.alpha.max() <= 1 + 1e-6


@settings(max_examples=10, **COMMON)
@given(P=st.integers(2, 300), seed=st.integers(0, 10_000))
def test_oracle_permutation_invariance(P, seed):
    s, cam = scene(P, seed)
    perm = torch.randperm(P, generator=torch.Generator().manual_seed(seed))
    a = oracle_render(s, cam, (1, 1, 1))
    b = oracle_render(tuple(t[perm] for t in s), cam, (1, 1, 1))
    assert np.array_equal(a.radii[perm.numpy()], b.radii)
    # same per-pixel sequences (depth order) => identical arithmetic, identical images
    assert np.array_equal(a.color, b.color) and np.array_equal(a.alpha, b.alpha) and np.array_equal(a.n_contrib, b.n_contrib)


@settings(max_examples=10, **COMMON)
@given(P=st.integers(1, 300), seed=st.integers(0, 10_000), k=st.floats(1.0, 3.0))
def test_oracle_radii_and_rects_grow_with_scale(P, seed, k):
    s, cam = scene(P, seed)
    a = oracle_render(s, cam, (1, 1, 1))
    b = oracle_render((s[0], s[1] * np.float32(k), s[2], s[3], s[4]), cam, (1, 1, 1))
    live = a.radii > 0
    assert (b.radii[live] >= a.radii[live]).all()
    assert (b.tiles_touched[live] >= a.tiles_touched[live]).all()


# ---- the same invariants on the CUDA path ---------------------------------------------------------------------------
def cuda_render(s, cam, bg, cols=None, grads=False):
    from dreammesh4d_b200 import rasterizer as R
    means, scales, rots, opac, c = s
    dev = "cuda"
    t = [x.to(dev).requires_grad_(grads) for x in (means, scales, rots, opac, c if cols is None else cols)]
    vp = R.make_view_params(cam[0][None].to(dev), cam[1][None].to(dev), torch.zeros(1, 3, device=dev), torch.tensor([cam[2]]),
                            torch.tensor([cam[3]]), torch.tensor(bg, dtype=torch.float32)[None])
    st_ = []
    color, radii, depth, alpha = R.rasterize_batch(t[0], t[3], t[1], t[2], t[4], vp, H, W, state_out=st_)
    return color[0], radii[0], depth[0], alpha[0], t, st_[0]


@pytest.mark.gpu
@settings(max_examples=8, **COMMON)
@given(P=st.integers(1, 2000), seed=st.integers(0, 10_000), r=st.floats(0, 1), g=st.floats(0, 1), b=st.floats(0, 1))
def test_cuda_partition_of_unity(P, seed, r, g, b):
    s, cam = scene(P, seed)
    col = torch.tensor([r, g, b], dtype=torch.float32)
    color, _, _, alpha, _, _ = cuda_render(s, cam, col.tolist(), cols=col.expand(P, 3).contiguous())
    assert float((color - col.to(color)[:, None, None]).abs().max()) <= 2e-6
    assert float(alpha.min()) >= 0 and float(alpha.max()) <= 1 + 1e-6


@pytest.mark.gpu
@settings(max_examples=8, **COMMON)
@given(P=st.integers(2, 2000), seed=st.integers(0, 10_000))
def test_cuda_permutation_invariance_forward_and_backward(P, seed):
    s, cam = scene(P, seed)
    perm = torch.randperm(P, generator=torch.Generator().manual_seed(seed))
    ca, ra, da, aa, ta, _ = cuda_render(s, cam, (1.0, 1.0, 1.0), grads=True)
    cb, rb, db, ab, tb, _ = cuda_render(tuple(t[perm] for t in s), cam, (1.0, 1.0, 1.0), grads=True)
    assert torch.equal(ra[perm.cuda()], rb) and torch.equal(ca, cb) and torch.equal(aa, ab) and torch.equal(da, db)
    g = torch.randn(3, H, W, generator=torch.Generator().manual_seed(seed + 7)).cuda()
    (ca * g).sum().backward()
    (cb * g).sum().backward()
    for x, y in zip(ta, tb):       # atomics: summation order only
        assert Hh.rel_linf(x.grad[perm.cuda()].cpu().numpy(), y.grad.cpu().numpy()) <= 2e-4


@pytest.mark.gpu
@settings(max_examples=6, **COMMON)
@given(P=st.integers(1, 1500), seed=st.integers(0, 10_000))
def test_cuda_zero_opacity_is_background_and_backward_is_linear(P, seed):
    s, cam = scene(P, seed)
    bg = (0.2, 0.5, 0.9)
    color, _, _, alpha, t, _ = cuda_render((s[0], s[1], s[2], torch.zeros_like(s[3]), s[4]), cam, bg, grads=True)
    assert float(alpha.abs().max()) == 0.0 and torch.equal(color, torch.tensor(bg).cuda()[:, None, None].expand_as(color))
    color.sum().backward()
    assert all(float(x.grad.abs().max()) == 0.0 for x in t)
    c2, _, d2, a2, t2, _ = cuda_render(s, cam, bg, grads=True)
    gen = torch.Generator().manual_seed(seed)
    g1, g2 = torch.randn(3, H, W, generator=gen).cuda(), torch.randn(3, H, W, generator=gen).cuda()
    ga = torch.autograd.grad((c2 * g1).sum(), t2, retain_graph=True)
    gb = torch.autograd.grad((c2 * g2).sum(), t2, retain_graph=True)
    gc = torch.autograd.grad((c2 * (2.0 * g1 - 3.0 * g2)).sum(), t2)
    for a, b, c in zip(ga, gb, gc):
        want = 2.0 * a.double() - 3.0 * b.double()
        assert float((c.double() - want).abs().max()) <= 3e-4 * float(want.abs().max().clamp_min(1e-20))
