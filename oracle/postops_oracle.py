"""CPU oracle of the per-view image post-ops that follow the rasterizer.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and nothing else; the product path
(dreammesh4d_b200/postops.py -> libdm4d.so) never touches it.

Plain torch restatement, statement by statement, of
  custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:25-54 (Depth2Normal),
  :180-193 (mask, masked depth, xyz map, normal from distance), :212-218 (rendered normals), :229 (clamp)
and of the static twin custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_normal.py:172-206, which
differs in ONE way that matters for gradients: it detaches the depth outside the mask AFTER the position map was
built from it, so the stencil gradient reaches unmasked neighbours there (``static=True``).
Pinned: tests/golden/postops.npz holds inputs, outputs and autograd gradients produced by executing the
reference's own statements (tests/golden/make_postops_golden.py); tests/test_postops.py checks this file against them.
Autograd supplies the gradients, dtype follows the inputs (fp64 in the parity tests).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def depth2normal(x: torch.Tensor) -> torch.Tensor:
    """Depth2Normal.forward (:44-54): two zero-padded 3x3 central-difference convolutions on [B,3,H,W], -cross."""
    B, C, H, W = x.shape
    kx = torch.tensor([[0.0, 0.0, 0.0], [-1.0, 0.0, 1.0], [0.0, 0.0, 0.0]], dtype=x.dtype).view(1, 1, 3, 3)
    ky = torch.tensor([[0.0, -1.0, 0.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]], dtype=x.dtype).view(1, 1, 3, 3)
    dzdx = F.conv2d(x.reshape(B * C, 1, H, W), kx, padding=1).reshape(B, C, H, W)
    dzdy = F.conv2d(x.reshape(B * C, 1, H, W), ky, padding=1).reshape(B, C, H, W)
    return -torch.cross(dzdx, dzdy, dim=1)


def _detach_outside(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x[~mask] = x[~mask].detach()"""
    return torch.where(mask, x, x.detach())


def post_ops(rgb, normal_img, depth, alpha, rays_o, rays_d, static: bool = False, compute_normal_from_dist: bool = True):
    """One view.  rgb, normal_img [3,H,W]; depth, alpha [1,H,W]; rays_o, rays_d [H,W,3].
    Returns dict(render, normal, normal_from_dist | None, depth, mask) with the reference's shapes ([C,H,W])."""
    mask = alpha > 0.99
    mask3 = mask.repeat(3, 1, 1)
    depth_masked = _detach_outside(depth, mask)
    depth_for_xyz = depth if static else depth_masked
    out = {}
    if compute_normal_from_dist:
        xyz_map = rays_o + depth_for_xyz.permute(1, 2, 0) * rays_d
        nfd = depth2normal(xyz_map.permute(2, 0, 1).unsqueeze(0))[0]
        nfd = F.normalize(nfd, dim=0)
        out["normal_from_dist"] = _detach_outside(nfd * 0.5 * alpha + 0.5, mask3)
    else:
        out["normal_from_dist"] = None
    normal = F.normalize(normal_img, dim=0)
    out["normal"] = _detach_outside(normal * 0.5 * alpha + 0.5, mask3)
    out["render"] = rgb.clamp(0, 1)
    out["depth"] = depth_masked
    out["mask"] = alpha
    return out
