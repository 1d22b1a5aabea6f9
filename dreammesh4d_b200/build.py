"""Builds dreammesh4d_b200/lib/libdm4d.so (sm_100a only) with nvcc.

In-tree build: the .so travels with the repo snapshot to the GPU box; nothing is JIT-compiled at
import time.  ``python -m dreammesh4d_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
LIB = LIBDIR / "libdm4d.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# translation unit -> extra flags.  raster_preprocess.cu carries the bit-exact arithmetic
# spec (no FMA contraction); the render kernels are compiled with default contraction.
UNITS = {
    "raster_preprocess.cu": ["-fmad=false"],
    "raster_preprocess_bwd.cu": [],
    "raster_binning.cu": [],
    "raster_render.cu": [],
    "skin.cu": [],
    "postops.cu": [],
    "hexplane.cu": [],
    "graph.cu": [],
    "nhwc_norm.cu": [],
    "capi.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[Path]:
    return [CSRC / u for u in UNITS if (CSRC / u).exists()]


def source_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, flags)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dm4d.h"]
    for p in deps:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(repr((ARCH, COMMON, sorted(UNITS.items()), os.environ.get("DM4D_NVCC_EXTRA", ""))).encode())
    return h.hexdigest()


def is_stale() -> bool:
    """Stale = the .so was not built from the sources as they are now (content hash, not mtime: the library travels
    between machines with the repo snapshot)."""
    stamp = LIB.with_suffix(".so.srchash")
    if not LIB.exists() or not stamp.exists():
        return True
    return stamp.read_text().strip() != source_hash()


def build(force: bool = False, verbose: bool = False, out: Path | None = None, extra: str | None = None) -> Path:
    """``out`` / ``extra``: build a tuning variant (extra nvcc flags) into another file, leaving LIB untouched."""
    if out is None and not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    objdir = PKG / "build" / (out.stem if out is not None else "default")
    objdir.mkdir(parents=True, exist_ok=True)
    target = out if out is not None else LIB
    extra_flags = (extra if extra is not None else os.environ.get("DM4D_NVCC_EXTRA", "")).split()
    host_cc = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else None

    def compile_one(src: Path) -> Path:
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *ARCH, *COMMON, *UNITS[src.name], *extra_flags, "-c", str(src), "-o", str(obj)]
        if host_cc:
            cmd[1:1] = ["-ccbin", host_cc]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed for {src.name}")
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, *ARCH, "-shared", "-o", str(target), *map(str, objs)]
    if host_cc:
        cmd[1:1] = ["-ccbin", host_cc]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if out is None:
        LIB.with_suffix(".so.srchash").write_text(source_hash() + "\n")
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
