"""Per-view image post-ops: the oracle is pinned to the reference's own statements
(tests/golden/postops.npz, made by tests/golden/make_postops_golden.py); the fused CUDA kernels
(dm4d_postops_forward / _backward through dreammesh4d_b200.postops) are checked against the oracle and the golden
vectors.  Tolerances: forward 1e-4 relative L-inf (north-star, images), gradients 1e-3."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import postops_oracle as PO
from tests import helpers as Hh

GOLD = Path(__file__).resolve().parent / "golden" / "postops.npz"
KEYS = ("render", "normal", "normal_from_dist", "depth", "mask")


def _load(tag):
    z = np.load(GOLD)
    return {k[len(tag) + 1:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(tag + "_")}


@pytest.mark.parametrize("tag", ["temporal", "static"])
def test_postops_oracle_matches_reference_statements(tag):
    t = _load(tag)
    leaves = [t[k].clone().requires_grad_(True) for k in ("rgb", "nrm", "depth", "alpha")]
    out = PO.post_ops(*leaves, t["rays_o"], t["rays_d"], static=(tag == "static"))
    for k in KEYS:
        assert torch.allclose(out[k], t[f"out_{k}"], rtol=0, atol=1e-12), k
    loss = sum((out[k] * t[f"cot_{k}"]).sum() for k in KEYS)
    grads = torch.autograd.grad(loss, leaves)
    for k, g in zip(("rgb", "nrm", "depth", "alpha"), grads):
        assert torch.allclose(g, t[f"grad_{k}"], rtol=0, atol=1e-10), k
    # the two renderers really differ in the depth gradient (stencil reaches unmasked neighbours in the static one)
    other = PO.post_ops(*[l.detach().clone().requires_grad_(True) for l in leaves], t["rays_o"], t["rays_d"],
                        static=(tag != "static"))
    assert torch.allclose(other["normal_from_dist"], out["normal_from_dist"])


def _gpu_case(B, H, W, seed, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(B, 3, H, W, generator=g, dtype=dtype) * 1.4 - 0.2
    nrm = torch.randn(B, 3, H, W, generator=g, dtype=dtype)
    depth = 3.3 + 0.3 * torch.rand(B, 1, H, W, generator=g, dtype=dtype)
    a = torch.rand(B, 1, H, W, generator=g, dtype=dtype)
    alpha = torch.where(a > 0.35, 0.99 + 0.01 * a, a)
    alpha[:, :, : H // 4, : W // 3] = 0.0                               # an empty corner: depth 0, zero stencil normal
    depth = torch.where(alpha > 0, depth, torch.zeros_like(depth))
    rays_o = torch.randn(B, 1, 1, 3, generator=g, dtype=dtype).expand(B, H, W, 3).contiguous()
    rays_d = torch.nn.functional.normalize(torch.randn(B, H, W, 3, generator=g, dtype=dtype) * 0.1 +
                                           torch.tensor([0.0, 0.0, -1.0], dtype=dtype), dim=-1)
    return rgb, nrm, depth, alpha, rays_o, rays_d


@pytest.mark.gpu
@pytest.mark.parametrize("static", [False, True])
@pytest.mark.parametrize("B,H,W", [(2, 37, 53), (1, 128, 96)])
def test_postops_kernels_match_oracle(B, H, W, static):
    from dreammesh4d_b200 import postops
    rgb, nrm, depth, alpha, rays_o, rays_d = _gpu_case(B, H, W, seed=B * 100 + H)
    leaves = [x.clone().requires_grad_(True) for x in (rgb, nrm, depth, alpha)]
    outs = [PO.post_ops(leaves[0][b], leaves[1][b], leaves[2][b], leaves[3][b], rays_o[b], rays_d[b], static=static)
            for b in range(B)]
    ref = {k: torch.stack([o[k] for o in outs]).permute(0, 2, 3, 1) for k in KEYS}       # [B,H,W,C] like batch_forward
    g = torch.Generator().manual_seed(7)
    cot = {k: torch.randn(ref[k].shape, generator=g, dtype=torch.float64) for k in KEYS}
    ref_grads = torch.autograd.grad(sum((ref[k] * cot[k]).sum() for k in KEYS), leaves)

    d = lambda x: x.float().cuda()
    color6 = torch.cat([rgb, nrm], dim=1)
    c6, dp, al = d(color6).requires_grad_(True), d(depth).requires_grad_(True), d(alpha).requires_grad_(True)
    got = postops.post_ops(c6, dp, al, d(rays_o), d(rays_d), static=static)
    names = {"render": "comp_rgb", "normal": "comp_normal", "normal_from_dist": "comp_normal_from_dist",
             "depth": "comp_depth", "mask": "comp_mask"}
    for k in KEYS:
        assert got[names[k]].shape == ref[k].shape
        assert Hh.rel_linf(got[names[k]].detach().cpu().double(), ref[k].detach()) <= Hh.TOL_IMAGE, k
    loss = sum((got[names[k]] * d(cot[k])).sum() for k in KEYS)
    g6, gd, ga = torch.autograd.grad(loss, (c6, dp, al))
    assert Hh.rel_linf(g6[:, :3].cpu().double(), ref_grads[0]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(g6[:, 3:].cpu().double(), ref_grads[1]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(gd.cpu().double(), ref_grads[2]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(ga.cpu().double(), ref_grads[3]) <= Hh.TOL_GRAD


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["temporal", "static"])
def test_postops_kernels_on_golden_inputs(tag):
    from dreammesh4d_b200 import postops
    t = _load(tag)
    d = lambda x: x.float().cuda()
    c6 = d(torch.cat([t["rgb"], t["nrm"]], dim=0)[None]).requires_grad_(True)
    dp, al = d(t["depth"][None]).requires_grad_(True), d(t["alpha"][None]).requires_grad_(True)
    got = postops.post_ops(c6, dp, al, d(t["rays_o"][None]), d(t["rays_d"][None]), static=(tag == "static"))
    names = {"render": "comp_rgb", "normal": "comp_normal", "normal_from_dist": "comp_normal_from_dist",
             "depth": "comp_depth", "mask": "comp_mask"}
    loss = 0.0
    for k in KEYS:
        want = t[f"out_{k}"].permute(1, 2, 0)[None]
        assert Hh.rel_linf(got[names[k]].detach().cpu().double(), want) <= Hh.TOL_IMAGE, k
        loss = loss + (got[names[k]] * d(t[f"cot_{k}"].permute(1, 2, 0)[None])).sum()
    g6, gd, ga = torch.autograd.grad(loss, (c6, dp, al))
    assert Hh.rel_linf(g6[0, :3].cpu().double(), t["grad_rgb"]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(g6[0, 3:].cpu().double(), t["grad_nrm"]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(gd[0].cpu().double(), t["grad_depth"]) <= Hh.TOL_GRAD
    assert Hh.rel_linf(ga[0].cpu().double(), t["grad_alpha"]) <= Hh.TOL_GRAD


@pytest.mark.gpu
def test_postops_without_normal_from_dist_is_consistent():
    """compute_normal_from_dist=False (no rays in the batch): same images, and the same gradients as the full call
    when the normal-from-distance output receives no gradient."""
    from dreammesh4d_b200 import postops
    rgb, nrm, depth, alpha, rays_o, rays_d = _gpu_case(2, 40, 56, seed=11)
    d = lambda x: x.float().cuda()
    mk = lambda: (d(torch.cat([rgb, nrm], dim=1)).requires_grad_(True), d(depth).requires_grad_(True), d(alpha).requires_grad_(True))
    g = torch.Generator().manual_seed(3)
    cot = {k: torch.randn(2, 40, 56, c, generator=g).cuda() for k, c in
           (("comp_rgb", 3), ("comp_normal", 3), ("comp_depth", 1), ("comp_mask", 1))}
    a6, ad, aa = mk()
    full = postops.post_ops(a6, ad, aa, d(rays_o), d(rays_d))
    b6, bd, ba = mk()
    lite = postops.post_ops(b6, bd, ba, compute_normal_from_dist=False)
    assert "comp_normal_from_dist" not in lite
    for k in cot:
        assert torch.equal(full[k], lite[k]), k
    ga = torch.autograd.grad(sum((full[k] * cot[k]).sum() for k in cot), (a6, ad, aa))
    gb = torch.autograd.grad(sum((lite[k] * cot[k]).sum() for k in cot), (b6, bd, ba))
    for x, y in zip(ga, gb):
        assert torch.allclose(x, y, rtol=0, atol=1e-6 * float(x.abs().max()))
