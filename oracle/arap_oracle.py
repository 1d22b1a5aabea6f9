"""CPU ORACLE for the ARAP regulariser — TEST INFRASTRUCTURE ONLY.

Restates ARAPCoach (custom/threestudio-dreammesh4d/utils/arap_utils.py): one-ring neighbours :72-81, cotangent
weights, dense variant :100-175 (W[i,j] = 0.5*cot assigned per directed face edge, W + W^T), edge matrices :177-181
and compute_arap_energy with supplied rotations :183-224, in plain torch (any dtype, autograd).
tests/golden/arap.npz holds vectors made by executing the reference class itself (tests/golden/make_arap_golden.py).
Rotation matrices come from pypose's SO3.matrix() in the caller (system/sugar_4dgen.py:373-375); semantics restated
in oracle/skin_oracle.py (parity unpinned for that third-party op)."""
from __future__ import annotations

import torch

from .skin_oracle import q_act


def one_ring(faces: torch.Tensor, V: int):
    nb = [set() for _ in range(V)]
    for f in faces.tolist():
        for j in range(3):
            nb[f[j]].add(f[(j + 1) % 3]); nb[f[j]].add(f[(j + 2) % 3])
    return [sorted(s) for s in nb]


def cot_weight_matrix(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """Dense V x V weight matrix exactly as arap_utils.py:108-150 (small meshes only)."""
    V = verts.shape[0]
    fv = verts[faces]
    v0, v1, v2 = fv[:, 0], fv[:, 1], fv[:, 2]
    A, B, C = (v1 - v2).norm(dim=1), (v0 - v2).norm(dim=1), (v0 - v1).norm(dim=1)
    s = 0.5 * (A + B + C)
    area = (s * (s - A) * (s - B) * (s - C)).clamp(min=1e-12).sqrt()
    A2, B2, C2 = A * A, B * B, C * C
    cot = torch.stack([(B2 + C2 - A2) / area, (A2 + C2 - B2) / area, (A2 + B2 - C2) / area], dim=1) / 4.0
    W = torch.zeros(V, V, dtype=verts.dtype)
    i = faces[:, [0, 1, 2]].flatten()
    j = faces[:, [1, 2, 0]].flatten()
    W[i, j] = 0.5 * cot.flatten()
    return W + W.T


def arap_energy(rest: torch.Tensor, faces: torch.Tensor, verts_t: torch.Tensor, vert_rot_xyzw_t: torch.Tensor) -> torch.Tensor:
    """Energy per timestamp [T] for verts_t [T,V,3], rotations [T,V,4] xyzw."""
    V = rest.shape[0]
    W = cot_weight_matrix(rest, faces)
    nb = one_ring(faces, V)
    ii = torch.tensor([i for i in range(V) for _ in nb[i]])
    jj = torch.tensor([j for i in range(V) for j in nb[i]])
    w = W[ii, jj]
    e = rest[ii] - rest[jj]
    out = []
    for t in range(verts_t.shape[0]):
        ep = verts_t[t][ii] - verts_t[t][jj]
        s = ep - q_act(vert_rot_xyzw_t[t][ii], e)
        out.append((w * (s * s).sum(-1)).sum())
    return torch.stack(out)


def mesh_normal_consistency(verts_t: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """pytorch3d.loss.mesh_normal_consistency restated (PARITY UNPINNED: pytorch3d@stable is not installed; call site
    custom/threestudio-dreammesh4d/system/sugar_4dgen.py:214-225): for every pair of faces sharing an edge (v0,v1)
    with opposite vertices a, b: n0 = (v1-v0)x(a-v0), n1 = (v1-v0)x(b-v0), term = 1 - cosine_similarity(n0, -n1);
    mean over pairs per mesh, mean over the batch of meshes.  verts_t [T,V,3]."""
    from collections import defaultdict
    edge_faces = defaultdict(list)
    for fi, f in enumerate(faces.tolist()):
        for k in range(3):
            a, b = f[k], f[(k + 1) % 3]
            edge_faces[(min(a, b), max(a, b))].append(f[(k + 2) % 3])
    rows = []
    for (lo, hi), opp in edge_faces.items():
        for x in range(len(opp)):
            for y in range(x + 1, len(opp)):
                rows.append((lo, hi, opp[x], opp[y]))
    idx = torch.tensor(rows)
    v0, v1, a, b = (verts_t[:, idx[:, k]] for k in range(4))
    n0 = torch.cross(v1 - v0, a - v0, dim=-1)
    n1 = torch.cross(v1 - v0, b - v0, dim=-1)
    loss = 1 - torch.cosine_similarity(n0, -n1, dim=-1)
    return loss.mean(dim=1).mean()
