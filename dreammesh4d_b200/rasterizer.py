"""Host side of the rasterizer: autograd binding over the C ABI plus the drop-in
``diff_gaussian_rasterization`` surface.

Mirrors the Python API of the un-vendored ``diff-gaussian-rasterization`` (ashawkey fork) that
the reference imports at
custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py:8-11 and calls at
:129-144 (settings) and :169-178 / :202-211 (4-tuple return ``color, radii, depth, alpha``);
surface restated in SURVEY.md Appendix A.1.  PyTorch is used for device memory, streams and
autograd plumbing only; all arithmetic runs in libdm4d.so.
"""
from __future__ import annotations

import ctypes
import os
import warnings
import weakref
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import DM4D_VIEW_STRIDE, RasterDesc, check, ptr


# --------------------------------------------------------------------------------------------
# view parameter blocks
# --------------------------------------------------------------------------------------------
def make_view_params(viewmatrix: torch.Tensor, projmatrix: torch.Tensor, campos: torch.Tensor, tanfovx, tanfovy,
                     bg: torch.Tensor, scale_modifier: float = 1.0,
                     set_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Packs per-view camera parameters into the [B, 48] fp32 block of include/dm4d.h.

    ``viewmatrix`` / ``projmatrix`` are the transposed (row-vector) [B,4,4] matrices produced by
    threestudio/utils/ops.py:398-413.  Everything stays on the device (no host sync).
    """
    dev = viewmatrix.device
    B = viewmatrix.shape[0]
    vp = torch.zeros(B, DM4D_VIEW_STRIDE, dtype=torch.float32, device=dev)
    vp[:, 0:16] = viewmatrix.reshape(B, 16)
    vp[:, 16:32] = projmatrix.reshape(B, 16)
    vp[:, 32:35] = campos.reshape(B, 3)
    vp[:, _lib.VIEW_TANFOVX] = torch.as_tensor(tanfovx, dtype=torch.float32, device=dev)
    vp[:, _lib.VIEW_TANFOVY] = torch.as_tensor(tanfovy, dtype=torch.float32, device=dev)
    vp[:, _lib.VIEW_SCALE_MOD] = float(scale_modifier)
    if set_index is not None:
        vp[:, _lib.VIEW_SET] = set_index.to(device=dev, dtype=torch.float32)
    bg = bg.to(device=dev, dtype=torch.float32)
    C = bg.shape[-1]
    vp[:, _lib.VIEW_BG:_lib.VIEW_BG + C] = bg.reshape(-1, C)
    return vp


# --------------------------------------------------------------------------------------------
# autograd binding
# --------------------------------------------------------------------------------------------
def _prep(t: torch.Tensor, k: int, name: str):
    """Returns (contiguous fp32 tensor, set stride in floats, n_sets)."""
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if t.dim() == 2:
        if t.shape[1] != k:
            raise ValueError(f"{name} must be [P,{k}] or [S,P,{k}], got {tuple(t.shape)}")
        return t, 0, 1
    if t.dim() == 3 and t.shape[2] == k:
        if t.shape[0] == 1:          # a single set is shared by every view: stride 0, gradients summed over views
            return t, 0, 1
        return t, t.shape[1] * k, t.shape[0]
    raise ValueError(f"{name} must be [P,{k}] or [S,P,{k}], got {tuple(t.shape)}")


class RasterWorkspace:
    """Optional persistent scratch (geom / bin / img / bwd sections, grow-only) for callers that run
    forward -> backward -> forward -> ... in stream order, e.g. a training loop or ``HostStreamedRasterStep``:
    the GB-sized sections are then allocated once instead of once per call.  A workspace serves ONE forward at a
    time: its backward must run before the next forward that uses the same workspace."""

    def __init__(self):
        self._t = {}

    def get(self, name: str, nbytes: int, device) -> torch.Tensor:
        t = self._t.get(name)
        if t is None or t.numel() < nbytes or t.device != device:
            t = torch.empty(int(nbytes * 1.1) + 256, dtype=torch.uint8, device=device)
            self._t[name] = t
        return t


class RasterState:
    """Caller-owned scratch of one batch (kept alive for the backward)."""

    def __init__(self, desc: RasterDesc, keep: list, capacity: int):
        self.desc, self.keep, self.capacity = desc, keep, capacity
        self.radii: Optional[torch.Tensor] = None
        self.alpha_version = (0, 0, 0)
        self.pending = None          # deferred overflow verification of a speculative capacity (CapacityBook)
        self.overflowed = False

    def status(self) -> tuple[int, bool]:
        """(num_rendered, overflow) — synchronises the current stream."""
        n = ctypes.c_int64(0)
        o = ctypes.c_int32(0)
        check(_lib.lib().dm4d_raster_status(ctypes.byref(self.desc), ctypes.byref(n), ctypes.byref(o),
                                            torch.cuda.current_stream().cuda_stream), "dm4d_raster_status")
        return int(n.value), bool(o.value)

    def overflow_flag(self) -> torch.Tensor:
        """Device-side int32 view of the overflow flag of this batch (no synchronisation; CUDA-graph capturable).
        The bin workspace starts with the status header {uint64 num_rendered; uint32 overflow; uint32 pad}
        (include/dm4d.h, "status header")."""
        return self.keep[8][8:12].view(torch.int32)

    def num_rendered_tensor(self) -> torch.Tensor:
        """Device-side int64 view of the instance count R of this batch (no synchronisation)."""
        return self.keep[8][0:8].view(torch.int64)

    def export_view(self, view: int):
        """Integer binning state of one view (ranges [T,2], point_list [R_v], n_contrib [H,W])."""
        d = self.desc
        dev = self.keep[0].device
        gx, gy = (d.W + 15) // 16, (d.H + 15) // 16
        ranges = torch.zeros(gx * gy, 2, dtype=torch.int32, device=dev)
        pl = torch.zeros(max(self.capacity, 1), dtype=torch.int32, device=dev)
        nc = torch.zeros(d.H, d.W, dtype=torch.int32, device=dev)
        check(_lib.lib().dm4d_raster_export_state(ctypes.byref(d), view, ptr(ranges), ptr(pl), pl.numel(), ptr(nc),
                                                  torch.cuda.current_stream().cuda_stream), "dm4d_raster_export_state")
        r_view = int(ranges[:, 1].max().item()) if ranges.numel() else 0
        return ranges, pl[:r_view], nc


class _RasterizeBatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, opacities, scales, rotations, colors, colors2, view_params, H, W,
                capacity, distinct_sets, state_out, workspace, plan=None, cov3D=None):
        l = _lib.lib()
        dev = means3D.device
        if dev.type != "cuda":
            raise _lib.Dm4dError("dreammesh4d_b200 rasterizer needs CUDA tensors (there is no CPU path)")
        channels = 3 if colors2 is None else 6
        m, ms, s0 = _prep(means3D, 3, "means3D")
        if cov3D is not None:          # cov3D_precomp of the replaced module: scales / rotations are not read
            cv, cvs, s1 = _prep(cov3D, 6, "cov3D_precomp")
            sc, ss, ro, rs, s2 = None, 0, None, 0, 1
        else:
            cv, cvs = None, 0
            sc, ss, s1 = _prep(scales, 3, "scales")
            ro, rs, s2 = _prep(rotations, 4, "rotations")
        op, os_, s3 = _prep(opacities.reshape(*opacities.shape[:-1], 1) if opacities.dim() >= 2 else opacities.reshape(-1, 1),
                            1, "opacities")
        co, cs, s4 = _prep(colors, 3, "colors_precomp")
        c2, c2s, s5 = (None, 0, 1) if colors2 is None else _prep(colors2, 3, "colors2")
        P = m.shape[-2]
        for t, nm in ((sc, "scales"), (ro, "rotations"), (op, "opacities"), (co, "colors"), (cv, "cov3D_precomp")):
            if t is not None and t.shape[-2] != P:
                raise ValueError(f"{nm} has {t.shape[-2]} rows, means3D has {P}")
        n_sets = max(s0, s1, s2, s3, s4, s5)
        for s in (s0, s1, s2, s3, s4, s5):
            if s not in (1, n_sets):
                raise ValueError("attribute set counts disagree")
        vp = view_params.contiguous().float()
        B = vp.shape[0]
        stream = torch.cuda.current_stream().cuda_stream

        d = RasterDesc()
        d.P, d.H, d.W, d.n_views, d.n_sets, d.channels = P, int(H), int(W), B, n_sets, channels
        d.flags = _lib.RASTER_VIEWS_DISTINCT_SETS if (distinct_sets and n_sets == B) else 0
        d.means3D, d.means3D_stride = ptr(m), ms
        d.scales, d.scales_stride = ptr(sc), ss
        d.rotations, d.rotations_stride = ptr(ro), rs
        d.opacities, d.opacities_stride = ptr(op), os_
        d.colors, d.colors_stride = ptr(co), cs
        d.colors2, d.colors2_stride = ptr(c2), c2s
        d.cov3D, d.cov3D_stride = ptr(cv), cvs
        d.view_params = ptr(vp)

        def sizes(cap):
            g, b, i, w = (ctypes.c_uint64(0) for _ in range(4))
            check(l.dm4d_raster_workspace_bytes(P, int(H), int(W), B, channels, int(cap), ctypes.byref(g),
                                                ctypes.byref(b), ctypes.byref(i), ctypes.byref(w)),
                  "dm4d_raster_workspace_bytes")
            return g.value, b.value, i.value, w.value

        u8 = dict(dtype=torch.uint8, device=dev)
        scratch = (lambda name, n: torch.empty(n, **u8)) if workspace is None else (lambda name, n: workspace.get(name, n, dev))
        radii = None if plan is not None else torch.empty(B, P, dtype=torch.int32, device=dev)
        color = torch.empty(B, channels, H, W, dtype=torch.float32, device=dev)
        depth = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        alpha = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)

        if plan is not None:
            # second pass over a planned batch (same geometry, other features): projection, binning and the depth
            # sort of `plan` are re-used; only the features are re-bound onto the sorted stream (dm4d.h)
            pd = plan.desc
            if (pd.P, pd.H, pd.W, pd.n_views, pd.channels) != (P, int(H), int(W), B, channels):
                raise ValueError("plan does not match this call (P, H, W, views or channels differ)")
            capacity = plan.capacity
            gb, bb, ib, wb = sizes(capacity)
            geom, bin_, img = plan.keep[7], scratch("bin", bb), scratch("img", ib)
            d.geom, d.geom_bytes, d.img, d.img_bytes = ptr(geom), gb, ptr(img), ib
            d.bin, d.bin_bytes, d.bin_capacity = ptr(bin_), bb, capacity
            check(l.dm4d_raster_render_features(ctypes.byref(pd), ctypes.byref(d), ptr(color), ptr(depth), ptr(alpha),
                                                stream), "dm4d_raster_render_features")
            radii = plan.radii.detach()          # same storage, fresh tensor object
        elif capacity is None:
            # exact sizing with one host read-back of num_rendered, like the replaced rasterizer:
            # plan into a capacity-0 workspace to learn R, then run the forward at capacity R
            gb, bb, ib, wb = sizes(0)
            geom, img, bin0 = scratch("geom", gb), scratch("img", ib), torch.empty(bb, **u8)
            d.geom, d.geom_bytes, d.img, d.img_bytes = ptr(geom), gb, ptr(img), ib
            d.bin, d.bin_bytes, d.bin_capacity = ptr(bin0), bb, 0
            n = ctypes.c_int64(0)
            check(l.dm4d_raster_plan(ctypes.byref(d), ptr(radii), ctypes.byref(n), stream), "dm4d_raster_plan")
            capacity = int(n.value)
            _, bb, _, _ = sizes(capacity)
            bin_ = scratch("bin", bb)
            d.bin, d.bin_bytes, d.bin_capacity = ptr(bin_), bb, capacity
            check(l.dm4d_raster_forward(ctypes.byref(d), ptr(color), ptr(depth), ptr(alpha), ptr(radii), stream),
                  "dm4d_raster_forward")
        else:
            capacity = int(capacity)
            gb, bb, ib, wb = sizes(capacity)
            geom, bin_, img = scratch("geom", gb), scratch("bin", bb), scratch("img", ib)
            d.geom, d.geom_bytes, d.img, d.img_bytes = ptr(geom), gb, ptr(img), ib
            d.bin, d.bin_bytes, d.bin_capacity = ptr(bin_), bb, capacity
            check(l.dm4d_raster_forward(ctypes.byref(d), ptr(color), ptr(depth), ptr(alpha), ptr(radii), stream),
                  "dm4d_raster_forward")

        # output.detach(): same storage and version counter, but no reference back to this node (no ctx <-> output cycle,
        # so the workspaces are released by reference counting, not by the cyclic GC)
        state = RasterState(d, [m, sc, ro, op, co, c2, vp, geom, bin_, img, alpha.detach(), color.detach(), depth.detach()], capacity)
        state.radii = radii
        state.cov3D = cv
        state.alpha_version = (alpha._version, color._version, depth._version)   # the backward reads these very images
        ctx.state = state
        ctx.workspace = workspace
        ctx.shapes = (means3D.shape, None if scales is None else scales.shape, None if rotations is None else rotations.shape,
                      opacities.shape, colors.shape, None if colors2 is None else colors2.shape,
                      None if cov3D is None else cov3D.shape)
        ctx.has_means2D = means2D is not None
        ctx.mark_non_differentiable(radii)
        if state_out is not None:
            state_out.append(state)
        return color, radii, depth, alpha

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha):
        l = _lib.lib()
        st: RasterState = ctx.state
        if st.pending is not None:
            st.pending.resolve(st)               # the header copy finished long ago: no stall
        d = st.desc
        m, sc, ro, op, co, c2, vp, geom, bin_, img, alpha, color, depth = st.keep
        if (alpha._version, color._version, depth._version) != st.alpha_version:
            raise RuntimeError("an image returned by the rasterizer was modified in place before backward(); the backward "
                               "reads the final transmittance and the composited totals from the forward's outputs (the "
                               "replaced rasterizer reads alpha the same way) — clone before editing")
        dev = m.device
        stream = torch.cuda.current_stream().cuda_stream
        _, _, _, wb = (ctypes.c_uint64(0) for _ in range(4))
        g_, b_, i_ = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        check(l.dm4d_raster_workspace_bytes(d.P, d.H, d.W, d.n_views, d.channels, d.bin_capacity, ctypes.byref(g_),
                                            ctypes.byref(b_), ctypes.byref(i_), ctypes.byref(wb)),
              "dm4d_raster_workspace_bytes")
        bwd = torch.empty(wb.value, dtype=torch.uint8, device=dev) if ctx.workspace is None else \
            ctx.workspace.get("bwd", wb.value, dev)
        d.bwd, d.bwd_bytes = ptr(bwd), wb.value

        gC = g_color.contiguous().float()
        gD = None if g_depth is None else g_depth.contiguous().float()
        gA = None if g_alpha is None else g_alpha.contiguous().float()
        need = ctx.needs_input_grad
        f32 = dict(dtype=torch.float32, device=dev)
        d_means3D = torch.empty_like(m) if need[0] else None
        d_means2D = torch.empty(d.n_views, d.P, 3, **f32) if (ctx.has_means2D and need[1]) else None
        d_opac = torch.empty_like(op) if need[2] else None
        d_scales = torch.empty_like(sc) if (sc is not None and need[3]) else None
        d_rots = torch.empty_like(ro) if (ro is not None and need[4]) else None
        cv = getattr(st, "cov3D", None)
        d_cov = torch.empty_like(cv) if (cv is not None and need[15]) else None
        d.dL_dcov3D = ptr(d_cov)
        d_colors = torch.empty_like(co) if need[5] else None
        d_colors2 = torch.empty_like(c2) if (c2 is not None and need[6]) else None
        check(l.dm4d_raster_backward(ctypes.byref(d), ptr(color), ptr(depth), ptr(alpha), ptr(gC), ptr(gD), ptr(gA), ptr(d_means3D),
                                     ptr(d_means2D), ptr(d_colors), ptr(d_colors2), ptr(d_opac), ptr(d_scales),
                                     ptr(d_rots), stream), "dm4d_raster_backward")
        sh = ctx.shapes
        rs = lambda t, s: None if t is None else t.reshape(s)
        return (rs(d_means3D, sh[0]), d_means2D, rs(d_opac, sh[3]), rs(d_scales, sh[1]), rs(d_rots, sh[2]),
                rs(d_colors, sh[4]), None if d_colors2 is None else rs(d_colors2, sh[5]),
                None, None, None, None, None, None, None, None, rs(d_cov, sh[6]))


def rasterize_batch(means3D, opacities, scales, rotations, colors, view_params, H, W, colors2=None, means2D=None,
                    capacity: Optional[int] = None, distinct_sets: bool = False, state_out: Optional[list] = None,
                    workspace: Optional[RasterWorkspace] = None, plan: Optional[RasterState] = None, cov3D=None):
    """Rasterizes a batch of views in one launch sequence.

    Attributes are ``[P,k]`` (shared by all views) or ``[S,P,k]`` (one set per timestamp; each view picks
    its set through ``view_params[:, 38]``).  ``means2D`` ([B,P,3], zeros) only serves as the holder of
    the screen-space mean gradient, as in the reference (``viewspace_points``,
    diff_sugar_rasterizer_temporal.py:108-113).  ``capacity=None`` sizes the binning workspace exactly
    with one host read-back (the replaced rasterizer's behaviour); an integer keeps the call fully
    asynchronous (check ``state.status()`` for overflow).  ``workspace``: optional persistent scratch
    (``RasterWorkspace``) for strictly alternating forward/backward call sequences.  ``plan``: the ``RasterState`` of an
    earlier call on the SAME means / scales / rotations / opacities / cameras — projection, binning and the depth sort
    are re-used and only ``colors`` (/``colors2``) are re-bound (the renderer's normal pass, temporal.py:202-211).
    ``cov3D`` ([P,6] or [S,P,6]: xx, xy, xz, yy, yz, zz): precomputed 3D covariances instead of ``scales`` / ``rotations``
    (``cov3D_precomp`` of the replaced module; the scale modifier does not apply).
    Returns ``color [B,C,H,W], radii [B,P] int32, depth [B,1,H,W], alpha [B,1,H,W]``.
    """
    return _RasterizeBatch.apply(means3D, means2D, opacities, scales, rotations, colors, colors2, view_params,
                                 int(H), int(W), capacity, distinct_sets, state_out, workspace, plan, cov3D)


# --------------------------------------------------------------------------------------------
# drop-in diff_gaussian_rasterization surface
# --------------------------------------------------------------------------------------------
class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435)


def sh_to_rgb(shs: torch.Tensor, means3D: torch.Tensor, campos: torch.Tensor, degree: int) -> torch.Tensor:
    """View-dependent colour from real spherical harmonics, the convention of the replaced rasterizer's preprocess
    (SURVEY.md Appendix A.2 step 8): ``shs [P, >= (degree+1)^2, 3]``, direction ``(p - campos) / |p - campos|``,
    ``rgb = max(0, SH(dir) + 0.5)`` (the clamp passes no gradient where it acts).  Degree 0 is
    ``SH2RGB`` of the reference (geometry/gaussian_base.py:39-40) followed by the clamp."""
    if degree < 0 or degree > 3:
        raise ValueError("sh_degree must be 0..3")
    if shs.dim() != 3 or shs.shape[-1] != 3 or shs.shape[1] < (degree + 1) ** 2:
        raise ValueError(f"shs must be [P, >= {(degree + 1) ** 2}, 3], got {tuple(shs.shape)}")
    sh = shs.unbind(dim=1)
    res = _SH_C0 * sh[0]
    if degree > 0:
        d = means3D - campos.reshape(1, 3).to(means3D)
        d = d / d.norm(dim=-1, keepdim=True)
        x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
        res = res - _SH_C1 * y * sh[1] + _SH_C1 * z * sh[2] - _SH_C1 * x * sh[3]
        if degree > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = res + _SH_C2[0] * xy * sh[4] + _SH_C2[1] * yz * sh[5] + _SH_C2[2] * (2.0 * zz - xx - yy) * sh[6] + \
                _SH_C2[3] * xz * sh[7] + _SH_C2[4] * (xx - yy) * sh[8]
            if degree > 2:
                res = res + _SH_C3[0] * y * (3.0 * xx - yy) * sh[9] + _SH_C3[1] * xy * z * sh[10] + \
                    _SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[11] + _SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[12] + \
                    _SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[13] + _SH_C3[5] * z * (xx - yy) * sh[14] + \
                    _SH_C3[6] * x * (xx - 3.0 * yy) * sh[15]
    return torch.clamp_min(res + 0.5, 0.0)


class CapacityBook:
    """Grow-only binning capacities for the per-view drop-in calls, replacing the replaced module's ``num_rendered``
    read-back in the middle of every forward.

    The first call of a (P, H, W, channels) shape is sized exactly (one read-back, as upstream does on every call);
    later calls run fully asynchronously at ``margin x`` the largest instance count seen.  Every speculative call
    queues a 16-byte copy of the device status header (include/dm4d.h) to pinned memory; it is read when it has
    landed (next call / the call's own backward) to keep the high-water mark current.  If a call ever exceeded its
    capacity — the instance count would have to jump by more than the margin between consecutive calls — that call
    rendered background: a RuntimeWarning says so, its backward returns zero gradients, the capacity grows, and the
    next call of the shape is sized exactly again.  ``exact = True`` (or DM4D_EXACT_SIZING=1) restores one read-back
    per call."""

    SLOTS = 256

    def __init__(self, margin: float = 1.5, slack: int = 1 << 16):
        self.margin, self.slack = margin, slack
        self.exact = bool(int(os.environ.get("DM4D_EXACT_SIZING", "0")))
        self.high: dict = {}             # shape key -> largest num_rendered seen
        self.force_exact: set = set()
        self._host = None
        self._free: list = []
        self._queue: list = []           # (event, slot, key, weakref(state))

    def capacity(self, key) -> Optional[int]:
        """None = size this call exactly (read-back)."""
        self.poll()
        if self.exact or key not in self.high or key in self.force_exact:
            self.force_exact.discard(key)
            return None
        return int(self.high[key] * self.margin) + self.slack

    def note_exact(self, key, n: int) -> None:
        self.high[key] = max(self.high.get(key, 0), int(n))

    def track(self, key, state: RasterState) -> None:
        """Queues the asynchronous header read-back of a speculative call."""
        if self._host is None:
            self._host = torch.zeros(self.SLOTS, 2, dtype=torch.int64).pin_memory()
            self._free = list(range(self.SLOTS))
        if not self._free:
            self.poll(block=True)
        slot = self._free.pop()
        self._host[slot].copy_(state.keep[8][0:16].view(torch.int64), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ent = _Pending(self, ev, slot, key)
        state.pending = ent
        self._queue.append((ent, weakref.ref(state)))

    def poll(self, block: bool = False) -> None:
        keep = []
        for ent, ref in self._queue:
            if ent.done:
                continue
            if block or ent.event.query():
                ent.resolve(ref())
            else:
                keep.append((ent, ref))
        self._queue = keep


class _Pending:
    def __init__(self, book: CapacityBook, event, slot: int, key):
        self.book, self.event, self.slot, self.key, self.done = book, event, slot, key, False

    def resolve(self, state: Optional[RasterState]) -> None:
        if self.done:
            return
        self.event.synchronize()
        total, flags = (int(x) for x in self.book._host[self.slot])
        self.book._free.append(self.slot)
        self.done = True
        self.book.note_exact(self.key, total)
        if flags & 0xffffffff:
            self.book.force_exact.add(self.key)
            if state is not None:
                state.overflowed = True
            warnings.warn(f"dreammesh4d_b200 rasterizer: a drop-in call produced {total} instances, more than its "
                          "speculative capacity; that call rendered background and its backward returns zero "
                          "gradients. The capacity has been raised and the next call is sized exactly "
                          "(set DM4D_EXACT_SIZING=1 to size every call exactly).", RuntimeWarning, stacklevel=3)
        if state is not None:
            state.pending = None


CAPACITY_BOOK = CapacityBook()


class GaussianRasterizer(nn.Module):
    """Same constructor/forward signature and 4-tuple return as the replaced module.

    Zero-edit fast path for the reference's renderer, which builds one rasterizer per view and calls it twice with the
    same means3D / scales / rotations / opacities objects (RGB, then normals as colours; temporal.py:144,169-178,
    202-211): the first call's plan (projection, binning, depth sort) is kept on the instance and the second call only
    re-binds its colours (``dm4d_raster_render_features``).  A hit requires the SAME tensor objects at the same
    ``_version`` — a freed-and-reallocated tensor can never alias a stale plan."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings
        self._plan = None            # (RasterState, (weakref, version) x 4)
        self._vp = None

    def _view_params(self) -> torch.Tensor:
        """[1,48] camera block, built once per instance with two small device ops."""
        if self._vp is None:
            rs = self.raster_settings
            dev = rs.viewmatrix.device
            tail = torch.tensor([float(rs.tanfovx), float(rs.tanfovy), float(rs.scale_modifier), 0.0, 0.0],
                                dtype=torch.float32).to(dev, non_blocking=True)
            bg = rs.bg.reshape(-1).to(device=dev, dtype=torch.float32)[:3]
            pad = torch.zeros(DM4D_VIEW_STRIDE - 43, dtype=torch.float32, device=dev)
            self._vp = torch.cat([rs.viewmatrix.reshape(16).float(), rs.projmatrix.reshape(16).float(),
                                  rs.campos.reshape(3).float(), tail, bg, pad]).reshape(1, DM4D_VIEW_STRIDE)
        return self._vp

    @staticmethod
    def _ident(*ts):
        return tuple((weakref.ref(t), t._version) for t in ts)

    def _plan_hit(self, *ts) -> Optional[RasterState]:
        if self._plan is None:
            return None
        state, ident = self._plan
        if state.overflowed or any(r() is not t or v != t._version for (r, v), t in zip(ident, ts)):
            return None
        return state

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Frustum test of the replaced module (``mark_visible`` -> in_frustum): view-space depth > 0.2 (SURVEY.md
        Appendix A.2 step 1).  Not called by the reference; part of the module's public surface."""
        with torch.no_grad():
            V = self.raster_settings.viewmatrix.to(positions)        # transposed (row-vector) world->view matrix
            return positions @ V[:3, 2] + V[3, 2] > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if shs is not None:
            # SH -> RGB as the replaced module does it in its preprocess (computeColorFromSH): evaluated here with device
            # tensor ops, then rasterized as precomputed colours.  Reached by the reference only in predict_step
            # (system/base.py:257, degree 0); autograd supplies the gradients incl. the view-direction term.
            colors_precomp = sh_to_rgb(shs, means3D, rs.campos, int(rs.sh_degree))
        vp = self._view_params()
        m2d = None if means2D is None else means2D.reshape(1, -1, 3)
        H, W = int(rs.image_height), int(rs.image_width)
        if cov3D_precomp is not None:
            # precomputed 3D covariances (never passed by the DreamMesh4D plugin, which always hands over scales / rotations:
            # diff_sugar_rasterizer_temporal.py:146-158): plain call, exact sizing, no plan re-use
            color, radii, depth, alpha = rasterize_batch(means3D, opacities, None, None, colors_precomp, vp, H, W,
                                                         means2D=m2d, cov3D=cov3D_precomp)
            return color[0], radii[0], depth[0], alpha[0]
        plan = self._plan_hit(means3D, scales, rotations, opacities)
        st: list = []
        if plan is not None:
            color, radii, depth, alpha = rasterize_batch(means3D, opacities, scales, rotations, colors_precomp, vp, H, W,
                                                         means2D=m2d, capacity=plan.capacity, state_out=st, plan=plan)
        else:
            key = (means3D.shape[-2], H, W, 3, means3D.device.index)
            cap = CAPACITY_BOOK.capacity(key)
            color, radii, depth, alpha = rasterize_batch(means3D, opacities, scales, rotations, colors_precomp, vp, H, W,
                                                         means2D=m2d, capacity=cap, state_out=st)
            if cap is None:
                CAPACITY_BOOK.note_exact(key, st[0].capacity)
            else:
                CAPACITY_BOOK.track(key, st[0])
            self._plan = (st[0], self._ident(means3D, scales, rotations, opacities))
        return color[0], radii[0], depth[0], alpha[0]
