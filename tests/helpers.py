"""Shared scene builders and comparison helpers for the parity tests."""
from __future__ import annotations

import math

import numpy as np
import torch

from dreammesh4d_b200 import synthetic
from dreammesh4d_b200.camera import get_cam_info_gaussian

TOL_IMAGE = 1e-4     # north-star: relative L-inf on rendered RGBA(+depth)
TOL_GRAD = 1e-3      # north-star: relative L-inf on gradients


def rel_linf(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(initial=0.0), 1e-30)
    return float(np.abs(a - b).max(initial=0.0) / denom)


def cameras(B: int, seed: int = 2, fovy_deg: float = 20.0, distance: float = 3.8):
    c2w, fovy = synthetic.random_orbit_cameras(B, seed=seed, distance=distance, fovy_deg=fovy_deg)
    V, PV, campos, tanx, tany = get_cam_info_gaussian(c2w, fovy, fovy)
    return V, PV, campos, tanx, tany


def random_scene(P: int, seed: int = 0, scale_lo: float = 0.005, scale_hi: float = 0.06):
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(P, 3, generator=g) - 0.5) * 0.9
    scales = torch.exp(torch.rand(P, 3, generator=g) * (math.log(scale_hi) - math.log(scale_lo)) + math.log(scale_lo))
    rots = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1)
    opac = torch.rand(P, 1, generator=g) * 0.94 + 0.05
    cols = torch.rand(P, 3, generator=g)
    return means, scales, rots, opac, cols


def sugar_gaussians(scene: synthetic.SugarScene):
    """Static-pose Gaussian attributes of a SugarScene computed with plain torch on the CPU
    (restating sugar.py:440-455,479-518,471-472,640-648) — test-side helper."""
    fv = scene.verts[scene.faces]                                   # [F,3,3]
    means = (fv[:, None] * scene.bary[None, :, :, None]).sum(dim=-2).reshape(-1, 3)
    n = torch.nn.functional.normalize(torch.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0], dim=-1), dim=-1)
    r1 = torch.nn.functional.normalize(fv[:, 0] - fv[:, 1], dim=-1)
    r2 = torch.nn.functional.normalize(torch.cross(n, r1, dim=-1), dim=-1)
    R = torch.stack([n, r1, r2], dim=-1)                             # columns
    R = R[:, None].expand(-1, scene.g, -1, -1).reshape(-1, 3, 3)
    quat = matrix_to_quaternion(R)
    scales = torch.cat([torch.full((means.shape[0], 1), scene.thickness), scene.log_scales.exp()], dim=-1)
    opac = torch.sigmoid(scene.densities)
    cols = scene.sh_dc[:, 0] * synthetic.C0 + 0.5
    normals = n[:, None].expand(-1, scene.g, -1).reshape(-1, 3)
    return means, scales, torch.nn.functional.normalize(quat, dim=-1), opac, cols, normals


def matrix_to_quaternion(R: torch.Tensor) -> torch.Tensor:
    """pytorch3d.transforms.matrix_to_quaternion semantics (wxyz)."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = R.reshape(-1, 9).unbind(-1)
    q_abs = torch.sqrt(torch.clamp(torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22,
                                                1 - m00 - m11 + m22], dim=-1), min=0))
    cand = torch.stack([
        torch.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], dim=-1)], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    best = q_abs.argmax(dim=-1)
    return cand[torch.arange(R.shape[0]), best]


def seeded_fill(module: torch.nn.Module, seed: int) -> None:
    """Deterministic parameter fill in state_dict order: matrices/filters ~ N(0, 0.5/sqrt(fan_in)), norm scales
    ~ 1 + N(0, 0.2), biases ~ N(0, 0.2).  Used by tests/golden/make_zero123_golden.py on the reference's modules and by
    the tests on the product's: equal names, order and shapes give equal weights."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.state_dict().items():
            if not p.dtype.is_floating_point:
                continue
            r = torch.randn(p.shape, generator=g)
            if p.dim() > 1:
                p.copy_(r * (0.5 / math.sqrt(p[0].numel())))
            elif name.endswith("weight"):
                p.copy_(1.0 + 0.2 * r)
            else:
                p.copy_(0.2 * r)


class SDSStubModel(torch.nn.Module):
    """Small seeded stand-in for the latent-diffusion network behind the SDS step (tests/golden/make_sds_golden.py and
    tests/test_sds.py build the SAME stub): ``moments`` plays the first-stage encoder + quant_conv (8x downsampling to 2x4
    moment channels), ``apply_model`` the hybrid-conditioned denoiser.  What it computes is irrelevant; that both
    sides compute the same thing is what lets the SDS arithmetic be compared."""

    scale_factor = 0.18215

    def __init__(self, seed: int = 3):
        super().__init__()
        self.enc = torch.nn.Conv2d(3, 8, 8, stride=8)
        self.cc_projection = torch.nn.Linear(772, 768)
        self.den = torch.nn.Conv2d(8, 4, 3, padding=1)
        self.den_t = torch.nn.Conv2d(8, 4, 1)
        self.ctx = torch.nn.Linear(768, 4)
        seeded_fill(self, seed)
        for p in self.parameters():
            p.requires_grad_(False)

    def moments(self, x):
        return self.enc(x)

    def apply_model(self, x, t, cond):
        xc = torch.cat([x] + list(cond["c_concat"]), dim=1)
        cc = torch.cat(list(cond["c_crossattn"]), dim=1)
        phase = torch.sin(t.float() / 1000.0 * 3.0).reshape(-1, 1, 1, 1)
        return self.den(xc) + phase * self.den_t(xc) + self.ctx(cc).mean(dim=1).reshape(-1, 4, 1, 1)


def sds_stub_inputs(seed: int, B: int = 3, n_frames: int = 5, hw: int = 48):
    g = torch.Generator().manual_seed(seed)
    return {"rgb": torch.rand(B, hw, hw, 3, generator=g),
            "elevation": torch.rand(B, generator=g) * 90 - 10, "azimuth": torch.rand(B, generator=g) * 360 - 180,
            "camera_distances": torch.full((B,), 3.8), "frame_indices": torch.randint(0, n_frames, (B,), generator=g),
            "c_crossattn": torch.randn(n_frames, 1, 768, generator=g), "c_concat": torch.randn(n_frames, 4, 32, 32, generator=g)}
