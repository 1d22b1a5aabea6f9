"""ctypes wrapper around oracle/raster_oracle.c — CPU ORACLE, TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see the header of raster_oracle.c): the reference's rasterizer is the
un-vendored ``diff-gaussian-rasterization`` (ashawkey fork, /root/reference/README.md:35)
and the reference ships no golden vectors for it.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference``
legs may import this module.  The product package (dreammesh4d_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"


def build(force: bool = False) -> None:
    """Compile the C oracle (both precisions) with the recipe in oracle/Makefile."""
    libs = [_BUILD / "liboracle_raster_f32.so", _BUILD / "liboracle_raster_f64.so"]
    src = _HERE / "raster_oracle.c"
    if not force and all(p.exists() and p.stat().st_mtime >= src.stat().st_mtime for p in libs):
        return
    subprocess.run(["make", "-C", str(_HERE), "-s", "-B" if force else "-s"], check=True)


_LIBS: dict[str, ctypes.CDLL] = {}


def _lib(precision: str) -> ctypes.CDLL:
    if precision not in _LIBS:
        build()
        lib = ctypes.CDLL(str(_BUILD / f"liboracle_raster_{precision}.so"))
        vp = ctypes.c_void_p
        lib.oracle_raster_create.restype = vp
        lib.oracle_raster_create.argtypes = [ctypes.c_int] * 4
        lib.oracle_raster_destroy.argtypes = [vp]
        lib.oracle_raster_set_ambiguity_margin.argtypes = [vp, ctypes.c_double]
        lib.oracle_raster_forward.argtypes = [vp] * 8 + [ctypes.c_double, ctypes.c_double, vp, ctypes.c_double]
        lib.oracle_raster_backward.argtypes = [vp] * 10
        lib.oracle_raster_num_rendered.restype = ctypes.c_int64
        lib.oracle_raster_num_rendered.argtypes = [vp]
        for name in ("color", "depth", "alpha", "radii", "rect", "tiles_touched", "ranges", "n_contrib",
                     "ambiguous", "xy", "gdepth", "conic_opacity"):
            fn = getattr(lib, f"oracle_raster_{name}")
            fn.restype = vp
            fn.argtypes = [vp]
        lib.oracle_raster_point_list.argtypes = [vp, vp, vp]
        _LIBS[precision] = lib
    return _LIBS[precision]


def _ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class RasterOracle:
    """One view of the rasterizer, forward + backward, on the CPU.

    ``precision`` is ``"f32"`` (the arithmetic the CUDA path must reproduce) or ``"f64"``.
    Matrices are the transposed (row-vector convention) 4x4s the plugin passes
    (threestudio/utils/ops.py:402-410).
    """

    def __init__(self, P: int, H: int, W: int, C: int = 3, precision: str = "f32"):
        self.P, self.H, self.W, self.C = P, H, W, C
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.lib = _lib(precision)
        self.h = self.lib.oracle_raster_create(P, H, W, C)
        if not self.h:
            raise ValueError("oracle_raster_create failed")
        self.gx, self.gy = (W + 15) // 16, (H + 15) // 16

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_raster_destroy(self.h)
            self.h = None

    def _a(self, x, shape):
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dtype).reshape(shape))
        return a

    def set_ambiguity_margin(self, rel: float) -> None:
        self.lib.oracle_raster_set_ambiguity_margin(self.h, float(rel))

    def forward(self, means3D, scales, rotations, opacities, colors, viewmatrix, projmatrix,
                tanfovx, tanfovy, bg, scale_modifier=1.0):
        P, C = self.P, self.C
        keep = [self._a(means3D, (P, 3)), self._a(scales, (P, 3)), self._a(rotations, (P, 4)),
                self._a(opacities, (P,)), self._a(colors, (P, C)), self._a(viewmatrix, (16,)),
                self._a(projmatrix, (16,)), self._a(bg, (C,))]
        rc = self.lib.oracle_raster_forward(self.h, *[_ptr(k) for k in keep[:7]], float(tanfovx), float(tanfovy),
                                            _ptr(keep[7]), float(scale_modifier))
        assert rc == 0
        return self.color, self.radii, self.depth, self.alpha

    def backward(self, dL_dcolor, dL_ddepth=None, dL_dalpha=None):
        P, C, H, W = self.P, self.C, self.H, self.W
        gC = self._a(dL_dcolor, (C, H, W))
        gD = None if dL_ddepth is None else self._a(dL_ddepth, (H, W))
        gA = None if dL_dalpha is None else self._a(dL_dalpha, (H, W))
        out = {
            "means3D": np.zeros((P, 3), self.dtype), "means2D": np.zeros((P, 3), self.dtype),
            "colors": np.zeros((P, C), self.dtype), "opacities": np.zeros((P, 1), self.dtype),
            "scales": np.zeros((P, 3), self.dtype), "rotations": np.zeros((P, 4), self.dtype),
        }
        rc = self.lib.oracle_raster_backward(self.h, _ptr(gC), _ptr(gD), _ptr(gA), _ptr(out["means3D"]),
                                             _ptr(out["means2D"]), _ptr(out["colors"]), _ptr(out["opacities"]),
                                             _ptr(out["scales"]), _ptr(out["rotations"]))
        assert rc == 0
        return out

    # ---- state accessors (copies) -------------------------------------------------------
    def _view(self, name, dtype, shape):
        p = getattr(self.lib, f"oracle_raster_{name}")(self.h)
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype)
        buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    @property
    def color(self): return self._view("color", self.dtype, (self.C, self.H, self.W))
    @property
    def depth(self): return self._view("depth", self.dtype, (1, self.H, self.W))
    @property
    def alpha(self): return self._view("alpha", self.dtype, (1, self.H, self.W))
    @property
    def radii(self): return self._view("radii", np.int32, (self.P,))
    @property
    def rect(self): return self._view("rect", np.int32, (self.P, 4))
    @property
    def tiles_touched(self): return self._view("tiles_touched", np.uint32, (self.P,))
    @property
    def ranges(self): return self._view("ranges", np.uint32, (self.gx * self.gy, 2))
    @property
    def n_contrib(self): return self._view("n_contrib", np.uint32, (self.H, self.W))
    @property
    def ambiguous(self): return self._view("ambiguous", np.uint8, (self.H, self.W)).astype(bool)
    @property
    def xy(self): return self._view("xy", self.dtype, (self.P, 2))
    @property
    def gaussian_depth(self): return self._view("gdepth", self.dtype, (self.P,))
    @property
    def conic_opacity(self): return self._view("conic_opacity", self.dtype, (self.P, 4))
    @property
    def num_rendered(self) -> int: return int(self.lib.oracle_raster_num_rendered(self.h))

    def point_list(self):
        """Sorted instance list: (gaussian ids, tile ids), each [R] uint32."""
        R = self.num_rendered
        ids = np.zeros(R, np.uint32)
        tiles = np.zeros(R, np.uint32)
        if R:
            self.lib.oracle_raster_point_list(self.h, _ptr(ids), _ptr(tiles))
        return ids, tiles


def cpu_threads() -> int:
    return int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
