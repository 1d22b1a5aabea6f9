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


# ---- mesh normal consistency (pytorch3d.loss.mesh_normal_consistency; sugar_4dgen.py:214-225) -------------------
def face_pairs(faces: torch.Tensor) -> torch.Tensor:
    """[n_pairs,4] int32 rows (v0, v1, a, b): every pair of faces sharing the undirected edge {v0,v1}, with the
    vertices a / b opposite to that edge in the two faces (setup, once; manifold edges give exactly one pair)."""
    f = faces.long().cpu()
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])              # directed edges, per face corner
    opp = torch.cat([f[:, 2], f[:, 0], f[:, 1]])
    lo, hi = e.min(dim=1).values, e.max(dim=1).values
    key = lo * (int(f.max()) + 1) + hi
    order = torch.argsort(key, stable=True)
    key, lo, hi, opp = key[order], lo[order], hi[order], opp[order]
    # runs of equal keys; manifold edges (runs of exactly two) are handled vectorised, the rest in a small loop
    change = torch.ones_like(key, dtype=torch.bool)
    change[1:] = key[1:] != key[:-1]
    run_start = torch.nonzero(change).flatten()
    run_len = torch.diff(torch.cat([run_start, torch.tensor([key.numel()])]))
    two = run_start[run_len == 2]
    out = [torch.stack([lo[two], hi[two], opp[two], opp[two + 1]], dim=1)]
    rows = []
    for st, ln in zip(run_start[run_len > 2].tolist(), run_len[run_len > 2].tolist()):
        for x in range(st, st + ln):
            for y in range(x + 1, st + ln):
                rows.append((int(lo[st]), int(hi[st]), int(opp[x]), int(opp[y])))
    if rows:
        out.append(torch.tensor(rows, dtype=torch.long))
    return torch.cat(out).to(dtype=torch.int32, device=faces.device).reshape(-1, 4)


class _NormalConsistencyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, pairs):
        l = _lib.lib()
        if verts.device.type != "cuda":
            raise _lib.Dm4dError("dreammesh4d_b200 needs CUDA tensors (there is no CPU path)")
        v = verts.detach().float().contiguous()
        T, V = v.shape[0], v.shape[1]
        loss = torch.empty(T, dtype=torch.float32, device=v.device)
        dv = torch.empty_like(v)
        check(l.dm4d_mesh_normal_consistency(ptr(pairs), pairs.shape[0], T, V, ptr(v), ptr(loss), ptr(dv),
                                             torch.cuda.current_stream().cuda_stream), "dm4d_mesh_normal_consistency")
        ctx.save_for_backward(dv)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dv,) = ctx.saved_tensors
        return dv * g[:, None, None], None


def mesh_normal_consistency(verts: torch.Tensor, pairs: torch.Tensor) -> torch.Tensor:
    """Scalar loss: mean over the T deformed meshes of the mean over face pairs of 1 - cos(n0, n1)."""
    return _NormalConsistencyFunction.apply(verts, pairs).mean()
