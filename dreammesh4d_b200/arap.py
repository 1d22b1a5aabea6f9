"""ARAP regulariser on the deformed mesh (host-side mirror of ARAPCoach,
custom/threestudio-dreammesh4d/utils/arap_utils.py:17-224, as used with supplied rotations by
system/sugar_4dgen.py:372-395).  Setup (one-ring + cotangent weights) is plain torch, once; the energy and its
gradient run in one fused kernel (dm4d_arap_energy) for all timestamps."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr


def cotangent_edge_weights(verts: torch.Tensor, faces: torch.Tensor):
    """One-ring CSR (row_ptr [V+1] int32, col [E] int32) and edge weights [E] with the reference's convention
    (arap_utils.py:100-175, dense variant): for face (v0,v1,v2) with side lengths A=|v1v2|, B=|v0v2|, C=|v0v1| and
    Heron area, the directed edges (v0->v1), (v1->v2), (v2->v0) get 0.5 * cot{a,b,c} / 4 with
    cota = (B^2+C^2-A^2)/area etc., and the weight of {i,j} is the sum of its two directed contributions.
    Avoids the reference's dense V x V matrix (10 GB at V = 50k)."""
    V = verts.shape[0]
    f = faces.long()
    fv = verts[f].double()
    v0, v1, v2 = fv[:, 0], fv[:, 1], fv[:, 2]
    A, B, C = (v1 - v2).norm(dim=1), (v0 - v2).norm(dim=1), (v0 - v1).norm(dim=1)
    s = 0.5 * (A + B + C)
    area = (s * (s - A) * (s - B) * (s - C)).clamp_(min=1e-12).sqrt()
    A2, B2, C2 = A * A, B * B, C * C
    cot = torch.stack([(B2 + C2 - A2) / area, (A2 + C2 - B2) / area, (A2 + B2 - C2) / area], dim=1) / 4.0
    i = f[:, [0, 1, 2]].reshape(-1)
    j = f[:, [1, 2, 0]].reshape(-1)
    val = 0.5 * cot.reshape(-1)
    # symmetric sum over the two directed copies of every undirected edge
    key = torch.cat([i * V + j, j * V + i])
    vals = torch.cat([val, val])
    ukey, inv = torch.unique(key, return_inverse=True)
    w = torch.zeros(ukey.shape[0], dtype=torch.float64, device=verts.device).index_add_(0, inv, vals)
    rows, cols = ukey // V, ukey % V            # sorted by row, then column
    counts = torch.bincount(rows, minlength=V)
    row_ptr = torch.zeros(V + 1, dtype=torch.int64, device=verts.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return row_ptr.int(), cols.int(), w.float()


class _ArapFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, vert_rot, rest_verts, row_ptr, col, w):
        l = _lib.lib()
        if verts.device.type != "cuda":
            raise _lib.Dm4dError("dreammesh4d_b200 needs CUDA tensors (there is no CPU path)")
        v = verts.detach().float().contiguous()
        q = vert_rot.detach().float().contiguous()
        T, V = v.shape[0], v.shape[1]
        energy = torch.empty(T, dtype=torch.float32, device=v.device)
        dv, dq = torch.empty_like(v), torch.empty_like(q)
        check(l.dm4d_arap_energy(ptr(rest_verts), ptr(row_ptr), ptr(col), ptr(w), T, V, ptr(v), ptr(q), ptr(energy),
                                 ptr(dv), ptr(dq), torch.cuda.current_stream().cuda_stream), "dm4d_arap_energy")
        ctx.save_for_backward(dv, dq)
        return energy

    @staticmethod
    def backward(ctx, g):
        dv, dq = ctx.saved_tensors
        return dv * g[:, None, None], dq * g[:, None, None], None, None, None, None


class ARAPEnergy:
    """``ARAPEnergy(rest_verts, faces)(verts [T,V,3], vert_rot [T,V,4] xyzw) -> energy [T]`` (differentiable)."""

    def __init__(self, rest_verts: torch.Tensor, faces: torch.Tensor):
        self.rest = rest_verts.detach().float().contiguous()
        self.row_ptr, self.col, self.w = cotangent_edge_weights(self.rest, faces)

    def __call__(self, verts: torch.Tensor, vert_rot: torch.Tensor) -> torch.Tensor:
        return _ArapFunction.apply(verts, vert_rot, self.rest, self.row_ptr, self.col, self.w)
