#!/usr/bin/env python
"""Builds tuning variants of libdm4d.so into dreammesh4d_b200/lib/variants/<name>.so (git-ignored, shipped by gpurun).
usage: python scripts/build_variants.py "w4=-DDM4D_RENDER_WARPS=4" "w2=-DDM4D_RENDER_WARPS=2" ..."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from dreammesh4d_b200 import build as b  # noqa: E402

out = b.LIBDIR / "variants"
out.mkdir(parents=True, exist_ok=True)
for old in out.glob("*.so"):
    old.unlink()
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    print(b.build(force=True, out=out / f"{name}.so", extra=flags))
