"""Deformation network: dreammesh4d_b200.deformation.HexPlaneDeformation against golden vectors made
by executing the reference's own DeformationNetwork (tests/golden/make_deformation_golden.py).  The reference's
state_dict loads with strict=True — parameter names, shapes and layouts are checkpoint-compatible
— and forward_dynamic_delta is reproduced for all timestamps in one batch (PyTorch lookup on the CPU; the fused CUDA
lookup is checked against the same statement in tests/test_hexplane.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from dreammesh4d_b200.deformation import HexPlaneDeformation

GOLD = Path(__file__).resolve().parent / "golden" / "deformation.npz"


def _load():
    z = np.load(GOLD)
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    t = {k: torch.from_numpy(z[k]) for k in z.files if not k.startswith("sd::")}
    return sd, t


def _net(fused):
    return HexPlaneDeformation(feat=8, base_res=(6, 7, 8, 5), multires=(1, 2), fused=fused)


def test_reference_checkpoint_loads_and_outputs_match():
    sd, t = _load()
    net = _net(fused=False)
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    dx, dr, ds, do = net(t["xyz"], t["ts"])
    for got, want in ((dx, t["dx"]), (dr, t["dr"]), (ds, t["ds"]), (do, t["do"])):
        assert got.shape == want.shape
        assert (got - want).abs().max() <= 1e-5 * want.abs().max()


@pytest.mark.gpu
def test_reference_checkpoint_outputs_match_with_the_fused_lookup():
    sd, t = _load()
    net = _net(fused=True)
    net.load_state_dict(sd, strict=True)
    net = net.cuda()
    outs = net(t["xyz"].cuda(), t["ts"].cuda())
    for got, want in zip(outs, (t["dx"], t["dr"], t["ds"], t["do"])):
        assert (got.cpu() - want).abs().max() <= 1e-4 * want.abs().max()
