"""Drop-in boundary (SURVEY.md §8b, B-inner): the module the reference imports as ``diff_gaussian_rasterization`` is
provided by dreammesh4d_b200/shims and accepts exactly the call shapes the reference's two renderers use
(tests/golden/api_calls.json, extracted from the reference by tests/golden/make_api_golden.py)."""
import importlib
import inspect
import json
import sys
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
    with pytest.raises(NotImplementedError):
        r(means3D=z(3), means2D=z(3), opacities=z(1), colors_precomp=z(3), cov3D_precomp=torch.zeros(4, 6))
