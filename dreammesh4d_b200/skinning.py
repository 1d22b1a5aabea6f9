"""Host side of the fused skinning + surface-bound Gaussian update (autograd binding over the C ABI).

Mirrors what DynamicSuGaRModel computes between the deformation network and the rasterizer:
``_get_timed_vertex_attributes_from_dg`` -> ``get_timed_gs_attributes`` -> ``get_timed_gs_normals``
(custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:487-613, 657-706, 357-364), for all
timestamps of a step in one launch sequence.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import SkinDesc, check, ptr

METHODS = {"lbs": 0, "dqs": 1, "hybrid": 2}


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


_INCIDENCE: dict = {}
USE_NODE_INCIDENCE = True     # node-centric vertex backward over incidence lists (node -> vertex slots): no reductions into
                              # the node tables; False: the C ABI's list-free backward (warp-aggregated reductions)
REPRODUCIBLE = False          # True: also gather the face -> vertex gradients through vertex -> corner lists instead of
                              # 16-byte vector reductions: the whole backward is bit-reproducible, ~10 % slower (C5: 672 vs
                              # 612 us at 8 timestamps)


def node_incidence(nbr_idx: torch.Tensor, M: int):
    """Incidence lists of the control nodes (``dm4d_skin_node_incidence``): ``inc_ptr`` [M+1], ``inc`` [V*K] int32.
    The deformation graph is fixed after start-up (dynamic_sugar.py:745-861), so the lists are built once per
    ``nbr_idx`` tensor (keyed on storage and version) and re-used by every backward."""
    key = (nbr_idx.data_ptr(), nbr_idx._version, tuple(nbr_idx.shape), M, nbr_idx.device)
    hit = _INCIDENCE.get(key)
    if hit is not None:
        return hit[0], hit[1]
    if torch.cuda.is_current_stream_capturing():
        raise _lib.Dm4dError("skinning: build the node incidence lists (one eager call) before capturing a CUDA graph")
    V, K = nbr_idx.shape
    dev = nbr_idx.device
    inc_ptr = torch.empty(M + 1, dtype=torch.int32, device=dev)
    inc = torch.empty(V * K, dtype=torch.int32, device=dev)
    scratch = torch.empty(M + 1, dtype=torch.int32, device=dev)
    check(_lib.lib().dm4d_skin_node_incidence(ptr(nbr_idx), V, K, M, ptr(inc_ptr), ptr(inc), ptr(scratch),
                                              torch.cuda.current_stream().cuda_stream), "dm4d_skin_node_incidence")
    if int(scratch[M]) != 0:
        raise ValueError(f"nbr_idx holds node indices outside [0, {M})")
    if len(_INCIDENCE) > 16:
        _INCIDENCE.clear()
    _INCIDENCE[key] = (inc_ptr, inc, nbr_idx)     # keeps nbr_idx alive: data_ptr stays unique
    return inc_ptr, inc


class _SkinFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node_trans, node_rot, node_scale, node_opacity, rest_verts, faces, nbr_idx, nbr_w, bary,
                rest_quat, method, want_normals):
        l = _lib.lib()
        dev = node_trans.device
        if dev.type != "cuda":
            raise _lib.Dm4dError("dreammesh4d_b200 skinning needs CUDA tensors (there is no CPU path)")
        if faces.dtype != torch.int32 or nbr_idx.dtype != torch.int32:
            raise TypeError("faces / nbr_idx must be int32 (convert once at setup)")
        # the reference hands out scale=None for dqs without d_scale and opacity=None unless hybrid
        # (dynamic_sugar.py:144-145,401-404); the C ABI takes NULL for the inputs a method does not read
        if node_scale is None and method != "dqs":
            raise ValueError(f"skinning_method={method!r} needs node_scale")
        if node_opacity is None and method == "hybrid":
            raise ValueError("skinning_method='hybrid' needs node_opacity")
        nt, nr = _f32(node_trans), _f32(node_rot)
        ns = None if node_scale is None else _f32(node_scale).reshape(*node_scale.shape[:2], 9)
        no = None if node_opacity is None else _f32(node_opacity).reshape(node_opacity.shape[0], -1)
        T, M = nt.shape[0], nt.shape[1]
        V, F, K, g = rest_verts.shape[0], faces.shape[0], nbr_idx.shape[1], bary.shape[0]
        P = F * g
        d = SkinDesc()
        d.n_t, d.V, d.F, d.M, d.K, d.g, d.method = T, V, F, M, K, g, METHODS[method]
        keep = [_f32(rest_verts), faces.contiguous(), nbr_idx.contiguous(), _f32(nbr_w), _f32(bary).reshape(g, 3),
                _f32(rest_quat), nt, nr, ns, no]
        (d.rest_verts, d.faces, d.nbr_idx, d.nbr_w, d.bary, d.rest_quat, d.node_trans, d.node_rot, d.node_scale,
         d.node_opacity) = [ptr(t) for t in keep]
        if USE_NODE_INCIDENCE and any(ctx.needs_input_grad[:4]):
            inc_ptr, inc = node_incidence(keep[2], M)
            d.node_inc_ptr, d.node_inc = ptr(inc_ptr), ptr(inc)
            keep += [inc_ptr, inc]
            if REPRODUCIBLE:
                vinc_ptr, vinc = node_incidence(keep[1], V)      # vertex -> face corners (same builder, K = 3)
                d.vert_inc_ptr, d.vert_inc = ptr(vinc_ptr), ptr(vinc)
                keep += [vinc_ptr, vinc]
        f32 = dict(dtype=torch.float32, device=dev)
        node_pre = torch.empty(T, M, 12, **f32)                  # per-(timestamp, node) pre-pass table
        d.node_scratch = ptr(node_pre)
        keep.append(node_pre)
        verts = torch.empty(T, V, 3, **f32)
        vert_rot = torch.empty(T, V, 4, **f32)
        means = torch.empty(T, P, 3, **f32)
        rots = torch.empty(T, P, 4, **f32)
        normals = torch.empty(T, P, 3, **f32) if want_normals else None
        check(l.dm4d_skin_forward(ctypes.byref(d), ptr(verts), ptr(vert_rot), ptr(means), ptr(rots), ptr(normals),
                                  torch.cuda.current_stream().cuda_stream), "dm4d_skin_forward")
        ctx.desc, ctx.keep = d, keep + [verts, vert_rot]
        ctx.shapes = (node_trans.shape, node_rot.shape, None if node_scale is None else node_scale.shape,
                      None if node_opacity is None else node_opacity.shape)
        ctx.want_normals = want_normals
        if not want_normals:
            normals = torch.empty(0, **f32)
            ctx.mark_non_differentiable(normals)
        return means, rots, normals, verts, vert_rot

    @staticmethod
    def backward(ctx, g_means, g_rots, g_normals, g_verts, g_vert_rot):
        l = _lib.lib()
        d: SkinDesc = ctx.desc
        verts, vert_rot = ctx.keep[-2], ctx.keep[-1]
        dev = verts.device
        f32 = dict(dtype=torch.float32, device=dev)
        c = lambda t: None if t is None else t.contiguous().float()
        g_means, g_rots, g_verts, g_vert_rot = c(g_means), c(g_rots), c(g_verts), c(g_vert_rot)
        g_normals = c(g_normals) if ctx.want_normals else None
        T, V, M = d.n_t, d.V, d.M
        dverts = torch.empty(T, V, 4, **f32)
        dvrot = torch.empty(T, V, 4, **f32)
        dn_t = torch.empty(T, M, 3, **f32)
        dn_r = torch.empty(T, M, 4, **f32)
        dn_s = torch.empty(T, M, 9, **f32)
        dn_o = torch.empty(T, M, **f32)
        if d.node_inc:
            scratch = torch.empty(T, V, 16, **f32)
            d.vert_scratch = ptr(scratch)
        if d.vert_inc:
            corner = torch.empty(T, d.F * 3, 8, **f32)
            d.corner_scratch = ptr(corner)
        check(l.dm4d_skin_backward(ctypes.byref(d), ptr(verts), ptr(vert_rot), ptr(g_means), ptr(g_rots),
                                   ptr(g_normals), ptr(g_verts), ptr(g_vert_rot), ptr(dverts), ptr(dvrot), ptr(dn_t),
                                   ptr(dn_r), ptr(dn_s), ptr(dn_o), torch.cuda.current_stream().cuda_stream),
              "dm4d_skin_backward")
        sh = ctx.shapes
        return (dn_t.reshape(sh[0]), dn_r.reshape(sh[1]), None if sh[2] is None else dn_s.reshape(sh[2]),
                None if sh[3] is None else dn_o.reshape(sh[3]), None, None, None, None, None, None, None, None)


def skin_gaussians(node_trans, node_rot, node_scale, node_opacity, rest_verts, faces, nbr_idx, nbr_w, bary, rest_quat,
                   method: str = "hybrid", want_normals: bool = True):
    """Deforms the mesh with the control-node attributes of ``T`` timestamps and updates the bound Gaussians.

    node_trans [T,M,3], node_rot [T,M,4] (xyzw, unit), node_scale [T,M,3,3], node_opacity [T,M,1] — the
    outputs of ``get_timed_dg_attributes`` (dynamic_sugar.py:367-405).  Static inputs: rest_verts [V,3],
    faces [F,3] int32, nbr_idx [V,K] int32, nbr_w [V,K], bary [g,3], rest_quat [P,4] wxyz.
    Returns means3D [T,P,3], rotations [T,P,4] (wxyz, normalised), normals [T,P,3] (or empty),
    verts [T,V,3], vert_rot [T,V,4] (xyzw).  Differentiable w.r.t. the four node tensors.
    """
    if method not in METHODS:
        raise ValueError(f"skinning_method must be one of {list(METHODS)}")
    return _SkinFunction.apply(node_trans, node_rot, node_scale, node_opacity, rest_verts, faces, nbr_idx, nbr_w,
                               bary, rest_quat, method, want_normals)


class _RestFramesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, complex_rot, faces, g):
        l = _lib.lib()
        dev = verts.device
        v, c, f = _f32(verts), _f32(complex_rot), faces.contiguous()
        P = f.shape[0] * g
        q = torch.empty(P, 4, dtype=torch.float32, device=dev)
        n = torch.empty(P, 3, dtype=torch.float32, device=dev)
        check(l.dm4d_sugar_rest_frames(ptr(v), ptr(f), ptr(c), v.shape[0], f.shape[0], g, ptr(q), ptr(n),
                                       torch.cuda.current_stream().cuda_stream), "dm4d_sugar_rest_frames")
        ctx.keep, ctx.g = (v, c, f), g
        return q, n

    @staticmethod
    def backward(ctx, g_q, g_n):
        l = _lib.lib()
        v, c, f = ctx.keep
        gq = None if g_q is None else g_q.contiguous().float()
        gn = None if g_n is None else g_n.contiguous().float()
        dv = torch.empty_like(v)
        dc = torch.empty_like(c)
        check(l.dm4d_sugar_rest_frames_backward(ptr(v), ptr(f), ptr(c), v.shape[0], f.shape[0], ctx.g, ptr(gq), ptr(gn),
                                                ptr(dv), ptr(dc), torch.cuda.current_stream().cuda_stream),
              "dm4d_sugar_rest_frames_backward")
        if gq is None:
            dc.zero_()
        return dv, dc, None, None


def sugar_rest_frames(verts: torch.Tensor, faces: torch.Tensor, complex_rot: torch.Tensor, g: int):
    """Rest-pose quaternions [P,4] (wxyz, normalised) and per-Gaussian unit face normals [P,3]
    (SuGaRModel.quaternions / get_gs_normals, sugar.py:490-526).  Differentiable w.r.t. ``verts`` and
    ``complex_rot`` (static stage); the static parameters are frozen in the dynamic stage
    (dynamic_sugar.py:79-87) and then no backward is ever launched."""
    if verts.device.type != "cuda":
        raise _lib.Dm4dError("dreammesh4d_b200 needs CUDA tensors (there is no CPU path)")
    if faces.dtype != torch.int32:
        raise TypeError("faces must be int32")
    return _RestFramesFunction.apply(verts, complex_rot, faces, g)
