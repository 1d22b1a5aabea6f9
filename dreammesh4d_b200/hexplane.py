"""Fused HexPlane multi-scale feature lookup (autograd binding over dm4d_hexplane_forward / _backward).

Mirror of interpolate_ms_features / HexPlaneField.forward
(custom/threestudio-dreammesh4d/geometry/deformation.py:141-174,242-248): one forward and one backward kernel for
all planes, scales and timestamps of a step instead of 24 grid_sample launches (+ products / views / cat) each way.
The plane parameters keep the reference's names, shapes and layout; their gradients are dense tensors as autograd's
grid_sample backward would produce (zeroed here, touched texels accumulated by the kernel).  CUDA only.
"""
from __future__ import annotations

import ctypes
import itertools
from typing import Sequence

import torch

from . import _lib
from ._lib import HexplaneDesc, check, ptr

PLANES = list(itertools.combinations(range(4), 2))      # (x,y) (x,z) (x,t) (y,z) (y,t) (z,t)


def _desc(coords: torch.Tensor, planes: Sequence[torch.Tensor], n_scales: int) -> HexplaneDesc:
    d = HexplaneDesc()
    feat = planes[0].shape[1]
    d.n_points, d.n_scales, d.feat = coords.shape[0], n_scales, feat
    d.coords = ptr(coords)
    for s in range(n_scales):
        res = [0, 0, 0, 0]
        for p, (i, j) in enumerate(PLANES):
            t = planes[s * 6 + p]
            if t.dim() != 4 or t.shape[0] != 1 or t.shape[1] != feat:
                raise ValueError(f"plane {s},{p}: expected [1,{feat},res_j,res_i], got {tuple(t.shape)}")
            for axis, size in ((i, t.shape[3]), (j, t.shape[2])):
                if res[axis] not in (0, size):
                    raise ValueError(f"scale {s}: planes disagree on the resolution of axis {axis}")
                res[axis] = size
            d.planes[s][p] = ptr(t)
        for k in range(4):
            d.res[s][k] = res[k]
    return d


class _HexPlaneFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, n_scales, *planes):
        if coords.device.type != "cuda":
            raise _lib.Dm4dError("dreammesh4d_b200 HexPlane lookup needs CUDA tensors (there is no CPU path)")
        if n_scales > _lib.HEX_MAX_SCALES or len(planes) != 6 * n_scales:
            raise ValueError(f"expected 6 planes for each of <= {_lib.HEX_MAX_SCALES} scales")
        coords = coords.detach().contiguous().float()
        planes_c = [p.detach().contiguous().float() for p in planes]
        d = _desc(coords, planes_c, n_scales)
        out = torch.empty(coords.shape[0], n_scales * d.feat, dtype=torch.float32, device=coords.device)
        check(_lib.lib().dm4d_hexplane_forward(ctypes.byref(d), ptr(out), torch.cuda.current_stream().cuda_stream),
              "dm4d_hexplane_forward")
        ctx.save_for_backward(coords, *planes_c)
        ctx.n_scales = n_scales
        return out

    @staticmethod
    def backward(ctx, g):
        coords, *planes = ctx.saved_tensors
        d = _desc(coords, planes, ctx.n_scales)
        need = ctx.needs_input_grad[2:]
        grads = [torch.zeros_like(p) if nd else None for p, nd in zip(planes, need)]
        arr = (ctypes.c_void_p * len(planes))(*[ptr(t) for t in grads])
        check(_lib.lib().dm4d_hexplane_backward(ctypes.byref(d), ptr(g.contiguous().float()), arr,
                                                torch.cuda.current_stream().cuda_stream), "dm4d_hexplane_backward")
        return (None, None, *grads)


def hexplane_features(coords4: torch.Tensor, grids: Sequence[Sequence[torch.Tensor]]) -> torch.Tensor:
    """``coords4 [N,4]``: normalised (x,y,z,t); ``grids[s][p]``: plane p of scale s, ``[1,F,res_j,res_i]``.
    Returns ``[N, S*F]`` — interpolate_ms_features(..., concat_features=True)."""
    flat = [p for planes in grids for p in planes]
    return _HexPlaneFeatures.apply(coords4, len(grids), *flat)
