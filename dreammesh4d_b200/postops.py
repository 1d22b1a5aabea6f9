"""Fused per-view image post-ops (autograd binding over dm4d_postops_forward / _backward).

Mirror of the tail of DiffGaussian.forward
(custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:180-193,212-218,229 and the static twin
diff_sugar_rasterizer_normal.py:172-206) plus the [B,H,W,C] stacking of GaussianBatchRenderer.batch_forward
(renderer/gaussian_batch_renderer.py:78-122): one forward kernel and two backward kernels for the whole batch
instead of ~30 launches (and two boolean-index host syncs) per view.  CUDA only.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import PostopsDesc, check, ptr


def _desc(color6, depth, alpha, rays_o, rays_d, flags) -> PostopsDesc:
    d = PostopsDesc()
    d.n_views, d.H, d.W, d.flags = color6.shape[0], color6.shape[2], color6.shape[3], flags
    d.color6, d.depth, d.alpha, d.rays_o, d.rays_d = ptr(color6), ptr(depth), ptr(alpha), ptr(rays_o), ptr(rays_d)
    return d


class _PostOps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, color6, depth, alpha, rays_o, rays_d, flags):
        if color6.device.type != "cuda":
            raise _lib.Dm4dError("dreammesh4d_b200 post-ops need CUDA tensors (there is no CPU path)")
        B, C, H, W = color6.shape
        if C != 6 or depth.shape != (B, 1, H, W) or alpha.shape != (B, 1, H, W):
            raise ValueError(f"expected color6 [B,6,H,W], depth/alpha [B,1,H,W]; got {tuple(color6.shape)}, "
                             f"{tuple(depth.shape)}, {tuple(alpha.shape)}")
        color6, depth, alpha = (t.contiguous().float() for t in (color6, depth, alpha))
        nfd = bool(flags & _lib.POSTOPS_NORMAL_FROM_DIST)
        if nfd:
            if rays_o.shape != (B, H, W, 3) or rays_d.shape != (B, H, W, 3):
                raise ValueError("rays_o / rays_d must be [B,H,W,3]")
            rays_o, rays_d = rays_o.contiguous().float(), rays_d.contiguous().float()
        else:
            rays_o = rays_d = None
        f32 = dict(dtype=torch.float32, device=color6.device)
        rgb, nrm = torch.empty(B, H, W, 3, **f32), torch.empty(B, H, W, 3, **f32)
        nfd_map = torch.empty(B, H, W, 3, **f32) if nfd else None
        dep, msk = torch.empty(B, H, W, 1, **f32), torch.empty(B, H, W, 1, **f32)
        d = _desc(color6, depth, alpha, rays_o, rays_d, flags)
        check(_lib.lib().dm4d_postops_forward(ctypes.byref(d), ptr(rgb), ptr(nrm), ptr(nfd_map), ptr(dep), ptr(msk),
                                              torch.cuda.current_stream().cuda_stream), "dm4d_postops_forward")
        ctx.save_for_backward(color6, depth, alpha, rays_o, rays_d)
        ctx.flags = flags
        if nfd:
            return rgb, nrm, nfd_map, dep, msk
        return rgb, nrm, dep, msk

    @staticmethod
    def backward(ctx, *grads):
        color6, depth, alpha, rays_o, rays_d = ctx.saved_tensors
        nfd = bool(ctx.flags & _lib.POSTOPS_NORMAL_FROM_DIST)
        if nfd:
            g_rgb, g_nrm, g_nfd, g_dep, g_msk = grads
        else:
            (g_rgb, g_nrm, g_dep, g_msk), g_nfd = grads, None
        c = lambda g: None if g is None else g.contiguous().float()
        g_rgb, g_nrm, g_nfd, g_dep, g_msk = c(g_rgb), c(g_nrm), c(g_nfd), c(g_dep), c(g_msk)
        B, _, H, W = color6.shape
        f32 = dict(dtype=torch.float32, device=color6.device)
        scratch = torch.empty(B, H, W, 6, **f32) if nfd else None
        d_color6, d_depth, d_alpha = torch.empty_like(color6), torch.empty_like(depth), torch.empty_like(alpha)
        d = _desc(color6, depth, alpha, rays_o, rays_d, ctx.flags)
        check(_lib.lib().dm4d_postops_backward(ctypes.byref(d), ptr(g_rgb), ptr(g_nrm), ptr(g_nfd), ptr(g_dep), ptr(g_msk),
                                               ptr(scratch), ptr(d_color6), ptr(d_depth), ptr(d_alpha),
                                               torch.cuda.current_stream().cuda_stream), "dm4d_postops_backward")
        return d_color6, d_depth, d_alpha, None, None, None


def post_ops(color6: torch.Tensor, depth: torch.Tensor, alpha: torch.Tensor, rays_o: Optional[torch.Tensor] = None,
             rays_d: Optional[torch.Tensor] = None, static: bool = False,
             compute_normal_from_dist: bool = True) -> Dict[str, torch.Tensor]:
    """``color6 [B,6,H,W]`` (rgb + rendered normals), ``depth``, ``alpha [B,1,H,W]`` as returned by the 6-channel
    rasterizer pass; ``rays_o``, ``rays_d [B,H,W,3]`` from the batch.  Returns the renderer's image outputs
    ``comp_rgb, comp_normal, comp_normal_from_dist [B,H,W,3]`` and ``comp_depth, comp_mask [B,H,W,1]``."""
    flags = (_lib.POSTOPS_STATIC if static else 0)
    if compute_normal_from_dist:
        if rays_o is None or rays_d is None:
            raise ValueError("compute_normal_from_dist needs rays_o and rays_d")
        flags |= _lib.POSTOPS_NORMAL_FROM_DIST
    outs = _PostOps.apply(color6, depth, alpha, rays_o, rays_d, flags)
    if compute_normal_from_dist:
        rgb, nrm, nfd, dep, msk = outs
        return {"comp_rgb": rgb, "comp_normal": nrm, "comp_normal_from_dist": nfd, "comp_depth": dep, "comp_mask": msk}
    rgb, nrm, dep, msk = outs
    return {"comp_rgb": rgb, "comp_normal": nrm, "comp_depth": dep, "comp_mask": msk}
