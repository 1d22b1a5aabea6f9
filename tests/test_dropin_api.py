"""Drop-in boundary (SURVEY.md §8b, B-inner): the module the reference imports as ``diff_gaussian_rasterization`` is
provided by dreammesh4d_b200/shims and accepts exactly the call shapes the reference's two renderers use
(tests/golden/api_calls.json, extracted from the reference by tests/golden/make_api_golden.py)."""
import importlib
import inspect
import json
from pathlib import Path

import pytest
import torch

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "api_calls.json").read_text())


def _shim():
    import dreammesh4d_b200
    dreammesh4d_b200.install_shim()
    return importlib.import_module("diff_gaussian_rasterization")


def test_shim_module_is_importable_under_the_reference_name():
    mod = _shim()
    assert Path(mod.__file__).resolve().parent.name == "diff_gaussian_rasterization"
    assert "shims" in str(Path(mod.__file__).resolve())
    for rec in GOLD.values():
        for name in rec["imports"]:
            assert hasattr(mod, name), name


@pytest.mark.parametrize("fname", sorted(GOLD))
def test_reference_call_shapes_are_accepted(fname):
    mod = _shim()
    rec = GOLD[fname]
    for kwargs in rec["settings_kwargs"]:
        assert list(mod.GaussianRasterizationSettings._fields) == kwargs          # same fields, same order (NamedTuple)
        mod.GaussianRasterizationSettings(**{k: None for k in kwargs})
    sig = inspect.signature(mod.GaussianRasterizer.forward)
    for kwargs in rec["forward_kwargs"]:
        sig.bind(None, **{k: None for k in kwargs})                                # raises TypeError on an unknown keyword
    assert set(rec["return_arity"]) == {4}


def test_argument_validation_matches_the_replaced_module():
    mod = _shim()
    s = mod.GaussianRasterizationSettings(8, 8, 0.2, 0.2, torch.ones(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3),
                                          False, False)
    r = mod.GaussianRasterizer(s)
    z = lambda k: torch.zeros(4, k)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=z(3), means2D=z(3), opacities=z(1), scales=z(3), rotations=z(4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=z(3), means2D=z(3), opacities=z(1), colors_precomp=z(3))
    # cov3D_precomp is accepted (CUDA tensors only: there is no CPU path)
    from dreammesh4d_b200._lib import Dm4dError
    with pytest.raises(Dm4dError, match="CUDA"):
        r(means3D=z(3), means2D=z(3), opacities=z(1), colors_precomp=z(3), cov3D_precomp=torch.zeros(4, 6))


def test_sh_to_rgb_degree0_is_the_reference_sh2rgb_and_bands_are_orthonormal():
    """``shs=`` path of the drop-in module (reached by the reference's predict_step with degree 0): degree 0 equals the
    reference's SH2RGB (geometry/gaussian_base.py:39-40: sh * 0.28209479177387814 + 0.5) with the rasterizer's clamp at 0;
    for degrees 1..3 the 16 basis functions are orthonormal on the sphere (constants and polynomials are consistent)."""
    import math
    from dreammesh4d_b200.rasterizer import sh_to_rgb
    g = torch.Generator().manual_seed(0)
    sh = torch.randn(50, 1, 3, generator=g, dtype=torch.float64)
    p = torch.randn(50, 3, generator=g, dtype=torch.float64)
    want = (sh[:, 0] * 0.28209479177387814 + 0.5).clamp_min(0)
    assert torch.allclose(sh_to_rgb(sh, p, torch.zeros(3, dtype=torch.float64), 0), want, atol=1e-15)
    # basis functions: unit coefficient on one band, colour - 0.5 without the clamp acting (scale the coefficient down)
    n = 200_000
    i = torch.arange(n, dtype=torch.float64) + 0.5
    phi, z = math.pi * (1 + 5 ** 0.5) * i, 1 - 2 * i / n                      # Fibonacci sphere: equal-area quadrature
    r = torch.sqrt(1 - z * z)
    dirs = torch.stack([r * torch.cos(phi), r * torch.sin(phi), z], dim=-1)
    Y = []
    for k in range(16):
        c = torch.zeros(n, 16, 3, dtype=torch.float64)
        c[:, k, :] = 0.1
        Y.append((sh_to_rgb(c, dirs, torch.zeros(3, dtype=torch.float64), 3)[:, 0] - 0.5) / 0.1)
    Y = torch.stack(Y, dim=1)                                                 # [n,16]
    gram = (Y.t() @ Y) * (4 * math.pi / n)
    assert (gram - torch.eye(16, dtype=torch.float64)).abs().max() < 2e-3
    # gradients flow to the coefficients and (degree > 0) to the positions; the clamp blocks them where it acts
    c = torch.randn(20, 16, 3, generator=g, dtype=torch.float64, requires_grad=True)
    q = torch.randn(20, 3, generator=g, dtype=torch.float64, requires_grad=True)
    out = sh_to_rgb(c, q, torch.tensor([0.1, -0.2, 3.0], dtype=torch.float64), 3)
    out.sum().backward()
    assert q.grad.abs().max() > 0 and c.grad.abs().max() > 0
    assert bool((c.grad[:, 0][out == 0] == 0).all())


def test_mark_visible_is_the_near_plane_test():
    from dreammesh4d_b200 import synthetic
    from dreammesh4d_b200.camera import get_cam_info_gaussian
    mod = _shim()
    c2w, fovy = synthetic.random_orbit_cameras(1, seed=4)
    V, PV, campos, tanx, tany = get_cam_info_gaussian(c2w, fovy, fovy)
    s = mod.GaussianRasterizationSettings(8, 8, float(tanx[0]), float(tany[0]), torch.ones(3), 1.0, V[0], PV[0], 0, campos[0], False, False)
    look = -campos[0] / campos[0].norm()
    pts = torch.stack([campos[0] + 0.1 * look, campos[0] + 0.3 * look, campos[0] - 1.0 * look, torch.zeros(3)])
    assert mod.GaussianRasterizer(s).markVisible(pts).tolist() == [False, True, False, True]
