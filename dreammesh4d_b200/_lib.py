"""ctypes binding of libdm4d.so (the C ABI declared in include/dm4d.h).

There is no CPU fallback: if the CUDA library is missing the import of any compute entry point
fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int32, c_int64, c_uint64, c_void_p
from pathlib import Path

PKG = Path(__file__).resolve().parent
# DM4D_LIB_PATH selects another build of the same library (kernel-tuning variants, scripts/tune_variants.sh)
LIB_PATH = Path(os.environ["DM4D_LIB_PATH"]) if os.environ.get("DM4D_LIB_PATH") else PKG / "lib" / "libdm4d.so"

DM4D_VIEW_STRIDE = 48
VIEW_TANFOVX, VIEW_TANFOVY, VIEW_SCALE_MOD, VIEW_SET, VIEW_BG = 35, 36, 37, 38, 40
RASTER_VIEWS_DISTINCT_SETS = 1


class RasterDesc(ctypes.Structure):
    """Mirror of ``dm4d_raster_desc`` (include/dm4d.h)."""
    _fields_ = [
        ("P", c_int32), ("H", c_int32), ("W", c_int32), ("n_views", c_int32), ("n_sets", c_int32),
        ("channels", c_int32), ("flags", c_int32), ("reserved1", c_int32),
        ("means3D", c_void_p), ("means3D_stride", c_int64),
        ("scales", c_void_p), ("scales_stride", c_int64),
        ("rotations", c_void_p), ("rotations_stride", c_int64),
        ("opacities", c_void_p), ("opacities_stride", c_int64),
        ("colors", c_void_p), ("colors_stride", c_int64),
        ("colors2", c_void_p), ("colors2_stride", c_int64),
        ("view_params", c_void_p),
        ("geom", c_void_p), ("geom_bytes", c_uint64),
        ("bin", c_void_p), ("bin_bytes", c_uint64),
        ("img", c_void_p), ("img_bytes", c_uint64),
        ("bwd", c_void_p), ("bwd_bytes", c_uint64),
        ("bin_capacity", c_int64),
        ("cov3D", c_void_p), ("cov3D_stride", c_int64), ("dL_dcov3D", c_void_p),
    ]


class SkinDesc(ctypes.Structure):
    """Mirror of ``dm4d_skin_desc`` (include/dm4d.h)."""
    _fields_ = [
        ("n_t", c_int32), ("V", c_int32), ("F", c_int32), ("M", c_int32), ("K", c_int32),
        ("g", c_int32), ("method", c_int32), ("reserved0", c_int32),
        ("rest_verts", c_void_p), ("faces", c_void_p), ("nbr_idx", c_void_p), ("nbr_w", c_void_p),
        ("bary", c_void_p), ("rest_quat", c_void_p), ("node_trans", c_void_p), ("node_rot", c_void_p),
        ("node_scale", c_void_p), ("node_opacity", c_void_p), ("node_inc_ptr", c_void_p), ("node_inc", c_void_p), ("vert_scratch", c_void_p),
        ("vert_inc_ptr", c_void_p), ("vert_inc", c_void_p), ("corner_scratch", c_void_p), ("node_scratch", c_void_p),
    ]


class PostopsDesc(ctypes.Structure):
    """Mirror of ``dm4d_postops_desc`` (include/dm4d.h)."""
    _fields_ = [("n_views", c_int32), ("H", c_int32), ("W", c_int32), ("flags", c_int32),
                ("color6", c_void_p), ("depth", c_void_p), ("alpha", c_void_p), ("rays_o", c_void_p), ("rays_d", c_void_p)]


POSTOPS_NORMAL_FROM_DIST, POSTOPS_STATIC = 1, 2
HEX_MAX_SCALES = 8


class HexplaneDesc(ctypes.Structure):
    """Mirror of ``dm4d_hexplane_desc`` (include/dm4d.h)."""
    _fields_ = [("n_points", c_int32), ("n_scales", c_int32), ("feat", c_int32), ("reserved", c_int32),
                ("coords", c_void_p), ("planes", (c_void_p * 6) * HEX_MAX_SCALES), ("res", (c_int32 * 4) * HEX_MAX_SCALES)]


# name -> (restype, argtypes); every symbol declared in include/dm4d.h
SIGNATURES = {
    "dm4d_raster_workspace_bytes": (ctypes.c_int, [c_int32] * 5 + [c_int64] + [POINTER(c_uint64)] * 4),
    "dm4d_raster_plan": (ctypes.c_int, [POINTER(RasterDesc), c_void_p, POINTER(c_int64), c_void_p]),
    "dm4d_raster_render": (ctypes.c_int, [POINTER(RasterDesc), c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_raster_forward": (ctypes.c_int, [POINTER(RasterDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_raster_render_features": (ctypes.c_int, [POINTER(RasterDesc), POINTER(RasterDesc), c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_raster_status": (ctypes.c_int, [POINTER(RasterDesc), POINTER(c_int64), POINTER(c_int32), c_void_p]),
    "dm4d_raster_backward": (ctypes.c_int, [POINTER(RasterDesc)] + [c_void_p] * 14),
    "dm4d_raster_export_state": (ctypes.c_int, [POINTER(RasterDesc), c_int32, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "dm4d_skin_forward": (ctypes.c_int, [POINTER(SkinDesc)] + [c_void_p] * 6),
    "dm4d_skin_backward": (ctypes.c_int, [POINTER(SkinDesc)] + [c_void_p] * 14),
    "dm4d_skin_node_incidence": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_sugar_rest_frames": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dm4d_sugar_rest_frames_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_arap_energy": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_mesh_normal_consistency": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dm4d_postops_forward": (ctypes.c_int, [POINTER(PostopsDesc)] + [c_void_p] * 6),
    "dm4d_postops_backward": (ctypes.c_int, [POINTER(PostopsDesc)] + [c_void_p] * 10),
    "dm4d_hexplane_forward": (ctypes.c_int, [POINTER(HexplaneDesc), c_void_p, c_void_p]),
    "dm4d_hexplane_backward": (ctypes.c_int, [POINTER(HexplaneDesc), c_void_p, POINTER(c_void_p), c_void_p]),
    "dm4d_graph_knn": (ctypes.c_int, [c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "dm4d_graph_geodesic_sweep": (ctypes.c_int, [c_int32, c_int32] + [c_void_p] * 9),
    "dm4d_groupnorm_nhwc_forward": (ctypes.c_int, [c_void_p] * 4 + [c_int32] * 4 + [ctypes.c_float] + [c_int32] * 2 + [c_void_p] * 4),
    "dm4d_groupnorm_nhwc_backward": (ctypes.c_int, [c_void_p] * 5 + [c_int32] * 4 + [ctypes.c_float] + [c_int32] * 2 + [c_void_p] * 4),
    "dm4d_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "dm4d_profile_collect": (ctypes.c_int, [POINTER(ctypes.c_double), POINTER(c_int64)]),
    "dm4d_kernel_name": (c_char_p, [ctypes.c_int]),
    "dm4d_last_error": (c_char_p, []),
    "dm4d_bias_residual_add_nhwc": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "dm4d_geglu": (ctypes.c_int, [c_void_p, ctypes.c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "dm4d_add_layernorm": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, ctypes.c_int64, c_int32, ctypes.c_float,
                                          c_int32, c_void_p, c_void_p, c_void_p]),
    "dm4d_version": (ctypes.c_int, []),
}

K_COUNT = 21

_lib = None


class Dm4dError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Loads libdm4d.so (once). Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise Dm4dError(
                f"{LIB_PATH} is missing: build it with `python -m dreammesh4d_b200.build` "
                "(or __graft_entry__.build()). dreammesh4d_b200 has no CPU / eager fallback.")
        l = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dm4d_last_error()
        raise Dm4dError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def profile_enable(on: bool) -> None:
    check(lib().dm4d_profile_enable(1 if on else 0), "dm4d_profile_enable")


def profile_collect() -> dict[str, tuple[float, int]]:
    """{kernel name: (total ms, launches)} since the last collect (synchronises the recorded events)."""
    ms = (ctypes.c_double * K_COUNT)()
    n = (c_int64 * K_COUNT)()
    check(lib().dm4d_profile_collect(ms, n), "dm4d_profile_collect")
    return {lib().dm4d_kernel_name(i).decode(): (ms[i], n[i]) for i in range(K_COUNT) if n[i]}
