"""ctypes wrapper + build of bench_ref_equiv/libdm4d_refequiv.so (see ref_equiv.cu).  Bench/test infrastructure:
the product package never imports this.  It links against libdm4d.so for the shared per-Gaussian projection math and
uses the product's descriptor / workspace layout, one view per call, like the rasterizer the reference binds."""
from __future__ import annotations

import ctypes
import subprocess
from ctypes import POINTER, c_int64, c_void_p
from pathlib import Path

import torch

from dreammesh4d_b200 import _lib
from dreammesh4d_b200 import build as product_build
from dreammesh4d_b200._lib import RasterDesc, ptr

HERE = Path(__file__).resolve().parent
SRC, LIB = HERE / "ref_equiv.cu", HERE / "libdm4d_refequiv.so"
_h = None


def build(force: bool = False) -> Path:
    product_build.build()
    deps = [SRC, product_build.LIB, product_build.CSRC / "raster_internal.cuh"]
    if not force and LIB.exists() and all(LIB.stat().st_mtime >= p.stat().st_mtime for p in deps):
        return LIB
    cmd = [product_build._nvcc(), "-ccbin", "/usr/bin/g++", *product_build.ARCH, *product_build.COMMON, "-shared", str(SRC),
           "-o", str(LIB), f"-L{product_build.LIBDIR}", "-ldm4d", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../dreammesh4d_b200/lib"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("ref_equiv build failed:\n" + r.stdout + r.stderr)
    return LIB


def lib() -> ctypes.CDLL:
    global _h
    if _h is None:
        if not LIB.exists():
            raise RuntimeError(f"{LIB} is missing: run bench_ref_equiv.ref_equiv.build() (or __graft_entry__.build())")
        _lib.lib()                      # libdm4d.so first (RTLD_GLOBAL not needed: resolved through DT_NEEDED + rpath)
        h = ctypes.CDLL(str(LIB))
        h.refeq_forward.restype = ctypes.c_int
        h.refeq_forward.argtypes = [POINTER(RasterDesc), c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_void_p]
        h.refeq_backward.restype = ctypes.c_int
        h.refeq_backward.argtypes = [POINTER(RasterDesc)] + [c_void_p] * 11
        _h = h
    return _h


class RefEquivView:
    """Forward + backward of ONE view with caller-side torch buffers (re-used across calls)."""

    def __init__(self, P: int, H: int, W: int, device):
        self.P, self.H, self.W, self.dev = P, H, W, device
        g, b, i, w = (ctypes.c_uint64(0) for _ in range(4))
        _lib.check(_lib.lib().dm4d_raster_workspace_bytes(P, H, W, 1, 3, 0, ctypes.byref(g), ctypes.byref(b), ctypes.byref(i),
                                                          ctypes.byref(w)), "workspace_bytes")
        u8 = dict(dtype=torch.uint8, device=device)
        self.ws = [torch.empty(x.value, **u8) for x in (g, b, i, w)]
        f32 = dict(dtype=torch.float32, device=device)
        self.color, self.depth, self.alpha = torch.empty(3, H, W, **f32), torch.empty(1, H, W, **f32), torch.empty(1, H, W, **f32)
        self.radii = torch.empty(P, dtype=torch.int32, device=device)
        self.grads = {"means3D": torch.empty(P, 3, **f32), "means2D": torch.empty(P, 3, **f32), "colors": torch.empty(P, 3, **f32),
                      "opacities": torch.empty(P, 1, **f32), "scales": torch.empty(P, 3, **f32), "rotations": torch.empty(P, 4, **f32)}

    def _desc(self, means, scales, rots, opac, cols, vp_row) -> RasterDesc:
        d = RasterDesc()
        d.P, d.H, d.W, d.n_views, d.n_sets, d.channels, d.flags = self.P, self.H, self.W, 1, 1, 3, 0
        d.means3D, d.scales, d.rotations, d.opacities, d.colors = ptr(means), ptr(scales), ptr(rots), ptr(opac), ptr(cols)
        d.view_params = ptr(vp_row)
        (d.geom, d.bin, d.img, d.bwd) = (ptr(t) for t in self.ws)
        d.geom_bytes, d.bin_bytes, d.img_bytes, d.bwd_bytes = (t.numel() for t in self.ws)
        d.bin_capacity = 0
        return d

    def forward_backward(self, means, scales, rots, opac, cols, vp_row, gC, gD=None, gA=None):
        """All tensors contiguous fp32 on the device; ``vp_row [48]`` with set index 0.  Returns num_rendered."""
        h = lib()
        s = torch.cuda.current_stream().cuda_stream
        d = self._desc(means, scales, rots, opac, cols, vp_row)
        n = c_int64(0)
        rc = h.refeq_forward(ctypes.byref(d), ptr(self.color), ptr(self.depth), ptr(self.alpha), ptr(self.radii), ctypes.byref(n), s)
        if rc:
            raise RuntimeError(f"refeq_forward failed ({rc}): {_lib.lib().dm4d_last_error().decode()}")
        G = self.grads
        rc = h.refeq_backward(ctypes.byref(d), ptr(self.alpha), ptr(gC), ptr(gD), ptr(gA), ptr(G["means3D"]), ptr(G["means2D"]),
                              ptr(G["colors"]), ptr(G["opacities"]), ptr(G["scales"]), ptr(G["rotations"]), s)
        if rc:
            raise RuntimeError(f"refeq_backward failed ({rc}): {_lib.lib().dm4d_last_error().decode()}")
        return int(n.value)
