"""Camera matrices for the Gaussian rasterizer — host-side mirror of
threestudio/utils/ops.py:359-413 (convert_pose, get_projection_matrix_gaussian,
get_cam_info_gaussian), batched over views and free of host synchronisation.
"""
from __future__ import annotations

import torch


def convert_pose(c2w: torch.Tensor) -> torch.Tensor:
    """ops.py:359-364 — flip the camera y and z axes (OpenGL -> COLMAP convention)."""
    # right-multiplication by diag(1,-1,-1,1) scales the columns; built from device-side ops only (no pageable
    # host->device copy, which would synchronise and cannot be captured into a CUDA graph)
    flip = torch.ones(4, dtype=c2w.dtype, device=c2w.device)
    flip[1:3] = -1.0
    return c2w * flip


def get_projection_matrix_gaussian(znear: float, zfar: float, tan_half_fovx: torch.Tensor,
                                   tan_half_fovy: torch.Tensor) -> torch.Tensor:
    """ops.py:367-387 — [B,4,4] (not yet transposed). right=-left, top=-bottom."""
    B = tan_half_fovx.shape[0]
    P = torch.zeros(B, 4, 4, dtype=tan_half_fovx.dtype, device=tan_half_fovx.device)
    top = tan_half_fovy * znear
    right = tan_half_fovx * znear
    P[:, 0, 0] = 2.0 * znear / (right - (-right))
    P[:, 1, 1] = 2.0 * znear / (top - (-top))
    P[:, 3, 2] = 1.0
    P[:, 2, 2] = zfar / (zfar - znear)
    P[:, 2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def get_cam_info_gaussian(c2w: torch.Tensor, fovx: torch.Tensor, fovy: torch.Tensor, znear: float = 0.1,
                          zfar: float = 100.0):
    """ops.py:398-413, batched: c2w [B,4,4], fov [B] (radians) ->
    (world_view_transform [B,4,4] transposed, full_proj_transform [B,4,4] transposed,
    camera_center [B,3], tanfovx [B], tanfovy [B])."""
    c2w = convert_pose(c2w.float())
    # inv_ex: same LU-based inverse as torch.inverse (ops.py:400) without the host-synchronising error check
    w2c = torch.linalg.inv_ex(c2w).inverse
    world_view = w2c.transpose(1, 2).contiguous()
    tanx, tany = torch.tan(fovx.float() * 0.5), torch.tan(fovy.float() * 0.5)
    proj = get_projection_matrix_gaussian(znear, zfar, tanx, tany).transpose(1, 2)
    full = torch.bmm(world_view, proj)
    center = torch.linalg.inv_ex(world_view).inverse[:, 3, :3]
    return world_view, full, center, tanx, tany
