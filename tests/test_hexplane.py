"""HexPlane multi-scale lookup: the oracle is pinned to the reference's own
HexPlaneField (tests/golden/hexplane.npz, made by tests/golden/make_hexplane_golden.py); the fused CUDA kernels
(dm4d_hexplane_forward / _backward) are checked against the oracle, the golden vectors and the PyTorch
(F.grid_sample) statement of the lookup.  Tolerances: 1e-4 relative L-inf forward, 1e-3 gradients."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import hexplane_oracle as HO
from tests import helpers as Hh

GOLD = Path(__file__).resolve().parent / "golden" / "hexplane.npz"


def _golden():
    z = np.load(GOLD)
    t = {k: torch.from_numpy(z[k]) for k in z.files}
    grids = [[t[f"plane_{s}_{p}"] for p in range(6)] for s in range(2)]
    grads = [[t[f"grad_{s}_{p}"] for p in range(6)] for s in range(2)]
    return t, grids, grads


def test_hexplane_oracle_matches_reference_code():
    t, grids, grads = _golden()
    leaves = [[p.clone().requires_grad_(True) for p in planes] for planes in grids]
    out = HO.hexplane_field(t["pts"], t["ts"], t["aabb"], leaves)
    assert torch.allclose(out, t["out"], rtol=0, atol=1e-12)
    got = torch.autograd.grad((out * t["cot"]).sum(), [p for planes in leaves for p in planes])
    for g, want in zip(got, [g for gs in grads for g in gs]):
        assert torch.allclose(g, want, rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_hexplane_kernels_on_golden_inputs():
    from dreammesh4d_b200.hexplane import hexplane_features
    t, grids, grads = _golden()
    d = lambda x: x.float().cuda()
    leaves = [[d(p).requires_grad_(True) for p in planes] for planes in grids]
    coords = torch.cat([HO.normalize_aabb(t["pts"], t["aabb"]), t["ts"]], dim=-1)
    out = hexplane_features(d(coords), leaves)
    assert Hh.rel_linf(out.detach().cpu().double(), t["out"]) <= Hh.TOL_IMAGE
    got = torch.autograd.grad((out * d(t["cot"])).sum(), [p for planes in leaves for p in planes])
    for g, want in zip(got, [g for gs in grads for g in gs]):
        assert g.shape == want.shape
        assert Hh.rel_linf(g.cpu().double(), want) <= Hh.TOL_GRAD


@pytest.mark.gpu
def test_hexplane_full_config_matches_oracle_and_grid_sample():
    """The YAML configuration: 32 features, base resolution [64,64,64,25], multires (1,2,4,8); 8 timestamps x 1000 nodes."""
    from dreammesh4d_b200.deformation import HexPlaneDeformation
    torch.manual_seed(0)
    net = HexPlaneDeformation().cuda()
    grid = net.deformation_net.grid
    with torch.no_grad():
        for planes in grid.grids:
            for p in planes:
                p.copy_(torch.rand_like(p) + 0.25)
    T, M = 8, 1000
    xyz = (torch.rand(M, 3, device="cuda") - 0.5) * 1.1
    pts = xyz.repeat(T, 1)
    ts = (torch.linspace(0, 1, T + 2, device="cuda")[1:-1] * 2 - 1).repeat_interleave(M)[:, None]
    cot = torch.randn(T * M, grid.feat_dim, device="cuda")
    params = [p for planes in grid.grids for p in planes]

    grid.fused = True
    out_f = grid(pts, ts)
    g_f = torch.autograd.grad((out_f * cot).sum(), params)
    grid.fused = False
    out_t = grid(pts, ts)
    g_t = torch.autograd.grad((out_t * cot).sum(), params)
    assert Hh.rel_linf(out_f.detach().cpu(), out_t.detach().cpu()) <= Hh.TOL_IMAGE
    for a, b in zip(g_f, g_t):
        assert Hh.rel_linf(a.cpu(), b.cpu()) <= Hh.TOL_GRAD

    # a slice against the fp64 oracle (CPU)
    sl = slice(0, 2 * M)
    coords = torch.cat([HO.normalize_aabb(pts[sl].cpu().double(), grid.aabb.detach().cpu().double()), ts[sl].cpu().double()], dim=-1)
    ref = HO.hexplane_features(coords, [[p.detach().cpu().double() for p in planes] for planes in grid.grids])
    assert Hh.rel_linf(out_f[sl].detach().cpu().double(), ref) <= Hh.TOL_IMAGE
