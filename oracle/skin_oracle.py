"""CPU ORACLE for sparse-control skinning + per-face surface-bound Gaussian update.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).

Restates, op for op, the reference's Python for this path in plain torch (any dtype, CPU, autograd):
  custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py
      :29-39   strain_tensor_to_matrix
      :487-613 _get_timed_vertex_attributes_from_dg  (LBS :530-549, DQS :551-564, hybrid :571-579,
                                                     rotation blend :582-586)
      :657-676 get_timed_gs_attributes, :726-743 _get_gs_xyz_from_vertex, :877-889 fuse_rotations,
      :347-364 get_timed_face_normals / get_timed_gs_normals
  custom/threestudio-dreammesh4d/utils/dual_quaternions.py:115-131,184-197,224-231,94-103
  custom/threestudio-dreammesh4d/geometry/sugar.py:440-455,471-472,479-518,521-526,640-648

PARITY UNPINNED for the third-party op semantics: pypose==0.6.7 (SO3 Log/Exp/mul/Act/matrix) and
pytorch3d@stable (Meshes.faces_normals_*, matrix_to_quaternion) are not installed here; their
published semantics are restated from SURVEY.md Appendix B.1-B.4.  tests/golden/ holds vectors made
by EXECUTING the reference's own functions above against these restated third-party ops
(tests/golden/make_skinning_golden.py), which pins the reference's composition of them.
Gradients: exact Euclidean autograd of these formulas (the reference mixes pypose's tangent-space
convention into the chain, SURVEY.md §7 H5 — not reproducible, documented deviation).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS_LIE = 1e-6


# ---- pypose SO3 semantics (xyzw), SURVEY.md Appendix B.1 -----------------------------------------
def q_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    av, aw = a[..., :3], a[..., 3:]
    bv, bw = b[..., :3], b[..., 3:]
    v = aw * bv + bw * av + torch.cross(av, bv, dim=-1)
    w = aw * bw - (av * bv).sum(-1, keepdim=True)
    return torch.cat([v, w], dim=-1)


def q_conj(q: torch.Tensor) -> torch.Tensor:
    return torch.cat([-q[..., :3], q[..., 3:]], dim=-1)


def q_act(q: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    v, w = q[..., :3], q[..., 3:]
    uv = torch.cross(v, p, dim=-1)
    return p + 2 * (w * uv + torch.cross(v, uv, dim=-1))


def so3_log(q: torch.Tensor) -> torch.Tensor:
    v, w = q[..., :3], q[..., 3:]
    n = v.norm(dim=-1, keepdim=True)
    big = n > EPS_LIE
    n_safe = torch.where(big, n, torch.ones_like(n))
    f_big = 2 * torch.atan(n_safe / w) / n_safe
    f_small = 2 / w - (2.0 / 3.0) * n * n / (w * w * w)
    return torch.where(big, f_big, f_small) * v


def so3_exp(x: torch.Tensor) -> torch.Tensor:
    th = x.norm(dim=-1, keepdim=True)
    big = th > EPS_LIE
    th_safe = torch.where(big, th, torch.ones_like(th))
    th2 = th * th
    a = torch.where(big, torch.sin(0.5 * th_safe) / th_safe, 0.5 - th2 / 48 + th2 * th2 / 3840)
    w = torch.where(big, torch.cos(0.5 * th_safe), 1 - th2 / 8 + th2 * th2 / 384)
    return torch.cat([a * x, w], dim=-1)


# ---- pytorch3d semantics, SURVEY.md Appendix B.4 --------------------------------------------------
def matrix_to_quaternion(R: torch.Tensor) -> torch.Tensor:
    """wxyz; candidate table + argmax selection."""
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
    return cand[torch.arange(R.shape[0]), best].reshape(*R.shape[:-2], 4)


def faces_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """Meshes.faces_normals_* then the plugin's extra F.normalize (sugar.py:495,522)."""
    fv = verts[..., faces, :]
    n = torch.cross(fv[..., 1, :] - fv[..., 0, :], fv[..., 2, :] - fv[..., 0, :], dim=-1)
    n = n / n.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    return F.normalize(n, dim=-1)


# ---- static SuGaR getters (sugar.py) --------------------------------------------------------------
def sugar_points(verts, faces, bary):
    fv = verts[faces]                                              # sugar.py:448
    pts = fv[:, None] * bary[None, :, :, None]                     # :451
    return pts.sum(dim=-2).reshape(-1, 3)                          # :453-455


def sugar_scaling(log_scales, thickness):
    return torch.cat([thickness * torch.ones(len(log_scales), 1, dtype=log_scales.dtype), torch.exp(log_scales)], dim=-1)


def sugar_quaternions(verts, faces, complex_rot, g):
    """sugar.py:490-518 -> [P,4] wxyz."""
    R0 = faces_normals(verts, faces)
    fv = verts[faces]
    b1 = F.normalize(fv[:, 0] - fv[:, 1], dim=-1)
    b2 = F.normalize(torch.cross(R0, b1, dim=-1))
    cn = F.normalize(complex_rot, dim=-1).view(len(faces), g, 2)
    R1 = cn[..., 0:1] * b1[:, None] + cn[..., 1:2] * b2[:, None]
    R2 = -cn[..., 1:2] * b1[:, None] + cn[..., 0:1] * b2[:, None]
    R = torch.cat([R0[:, None, :, None].expand(-1, g, -1, -1).clone(), R1[..., None], R2[..., None]], dim=-1).view(-1, 3, 3)
    return F.normalize(matrix_to_quaternion(R), dim=-1)


def sugar_opacity(densities):
    return torch.sigmoid(densities.view(-1, 1))


def sugar_points_rgb(sh_dc):
    return (sh_dc * 0.28209479177387814 + 0.5).view(-1, 3)


def strain_tensor_to_matrix(strain: torch.Tensor) -> torch.Tensor:
    """dynamic_sugar.py:29-39."""
    m = torch.zeros(*strain.shape[:-1], 9, dtype=strain.dtype)
    m[..., [0, 4, 8]] += 1.0
    m[..., [0, 4, 8]] += strain[..., :3]
    m[..., [1, 2, 5]] += strain[..., 3:]
    m[..., [3, 6, 7]] += strain[..., 3:]
    return m.reshape(*strain.shape[:-1], 3, 3)


# ---- vertex stage (dynamic_sugar.py:487-613) -------------------------------------------------------
def skin_vertices(rest_verts, nbr_idx, nbr_w, node_trans, node_rot, node_scale, node_opacity, method="hybrid"):
    """node_trans [T,M,3], node_rot [T,M,4] xyzw unit, node_scale [T,M,3,3], node_opacity [T,M,1].
    Returns verts [T,V,3], vert_rot [T,V,4] xyzw."""
    n_trans = node_trans[:, nbr_idx]                # [T,V,K,3]
    n_rot = node_rot[:, nbr_idx]                    # [T,V,K,4]
    w = nbr_w[None, :, :, None]
    x = rest_verts
    if method in ("lbs", "hybrid"):
        n_scale = node_scale[:, nbr_idx]            # [T,V,K,3,3]
        y = torch.matmul(n_scale, x[None, :, None, :, None]).squeeze(-1)          # :530-534  S x
        y = q_act(n_rot, y)                                                       # :535-537  R (S x)
        y = y + n_trans                                                           # :538
        x_lbs = (w * y).sum(dim=2)                                                # :548-549
    if method in ("dqs", "hybrid"):
        q_r = n_rot / n_rot.norm(dim=-1, keepdim=True)                            # dual_quaternions.py:123-124
        t0 = torch.cat([n_trans, torch.zeros_like(n_trans[..., :1])], dim=-1)
        q_d = q_mul(0.5 * t0, q_r)                                                # :126-130
        sq_r = (q_r * w).sum(dim=-2)                                              # dynamic_sugar.py:558
        sq_d = (q_d * w).sum(dim=-2)                                              # :559
        nrm = sq_r.norm(dim=-1, keepdim=True)                                     # dual_quaternions.py:194
        qn, dn = sq_r / nrm, sq_d / nrm                                           # :197
        trans = q_mul(2.0 * dn, q_conj(qn))[..., :3]                              # :230-231
        x_dqs = q_act(qn, x[None]) + trans                                        # :98-103
    if method == "lbs":
        xyz = x_lbs
    elif method == "dqs":
        xyz = x_dqs
    else:
        n_op = node_opacity[:, nbr_idx]                                           # :572
        lam = (nbr_w[None, ..., None] * n_op).sum(dim=-2)                         # :573-575
        lam = torch.clamp(lam + 0.4, max=1.0)                                     # :577
        xyz = lam * x_lbs + (1 - lam) * x_dqs                                     # :579
    rot = so3_exp((nbr_w[None, ..., None] * so3_log(n_rot)).sum(dim=-2))          # :583-586
    return xyz, rot


# ---- Gaussian stage (dynamic_sugar.py:657-676,726-743,877-889,347-364) ----------------------------
def gaussians_from_vertices(verts_t, vert_rot_t, faces, bary, rest_quat_wxyz):
    """verts_t [T,V,3], vert_rot_t [T,V,4] xyzw, bary [g,3], rest_quat [P,4] wxyz.
    Returns means3D [T,P,3], rotations [T,P,4] wxyz (normalised), normals [T,P,3]."""
    g = bary.shape[0]
    T = verts_t.shape[0]
    fv = verts_t[:, faces]                                                        # :733
    pts = (fv[:, :, None] * bary[None, None, :, :, None]).sum(dim=-2)             # :736-738
    means = pts.reshape(T, -1, 3)
    conn = faces.repeat_interleave(g, dim=0)                                      # :157-159
    bw = bary.repeat(faces.shape[0], 1)[..., None]                                # :154-156
    rl = so3_log(vert_rot_t[:, conn])                                             # :885
    dq = so3_exp((bw[None] * rl).sum(dim=-2))                                     # :887-888
    rest_xyzw = rest_quat_wxyz[None, :, [1, 2, 3, 0]]                             # :673
    q = q_mul(dq, rest_xyzw)[..., [3, 0, 1, 2]]                                   # :674-675
    rots = F.normalize(q, dim=-1)                                                 # :676
    normals = faces_normals(verts_t, faces).repeat_interleave(g, dim=1)           # :352-364
    return means, rots, normals


def deform_gaussians(scene, graph, node_trans, node_rot, node_scale, node_opacity, method="hybrid", dtype=torch.float32):
    """Full geometry path for a SugarScene-like object (fields verts, faces, bary, log_scales, complex_rot,
    densities, sh_dc, thickness, g) and a DeformGraph-like object (nbr_idx, nbr_w). Returns a dict of
    per-timestamp Gaussian sets as get_timed_gs_all_single_time would give per view (:708-724)."""
    c = lambda t: t.to(dtype)
    verts, faces = c(scene.verts), scene.faces
    rest_q = sugar_quaternions(verts, faces, c(scene.complex_rot), scene.g)
    xyz, rot = skin_vertices(verts, graph.nbr_idx, c(graph.nbr_w), c(node_trans), c(node_rot), c(node_scale),
                             c(node_opacity), method)
    means, rots, normals = gaussians_from_vertices(xyz, rot, faces, c(scene.bary), rest_q)
    return {"verts": xyz, "vert_rot": rot, "means3D": means, "rotations": rots, "normals": normals,
            "scales": sugar_scaling(c(scene.log_scales), scene.thickness), "opacities": sugar_opacity(c(scene.densities)),
            "colors": sugar_points_rgb(c(scene.sh_dc)), "rest_quat": rest_q}
