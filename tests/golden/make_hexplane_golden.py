#!/usr/bin/env python
"""Generates tests/golden/hexplane.npz by executing the reference's own HexPlaneField
(custom/threestudio-dreammesh4d/geometry/deformation.py: normalize_aabb, grid_sample_wrapper, init_grid_param,
interpolate_ms_features, class HexPlaneField — unmodified source, extracted by AST from /root/reference at generation
time only) on a small seeded configuration: inputs (points incl. some outside the box, timestamps), the randomised
planes, the [N, S*F] features and the autograd gradients w.r.t. every plane."""
import ast
import itertools
from pathlib import Path
from typing import Collection, Iterable, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/custom/threestudio-dreammesh4d/geometry/deformation.py")
WANT = {"normalize_aabb", "grid_sample_wrapper", "init_grid_param", "interpolate_ms_features", "HexPlaneField"}


def main():
    src = REF.read_text()
    tree = ast.parse(src)
    ns = {"torch": torch, "nn": nn, "F": F, "itertools": itertools, "Collection": Collection, "Iterable": Iterable,
          "Optional": Optional, "Sequence": Sequence}
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in WANT:
            exec(compile(ast.Module([node], []), str(REF), "exec"), ns)
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float64)
    field = ns["HexPlaneField"](1.0, {"grid_dimensions": 2, "input_coordinate_dim": 4, "output_coordinate_dim": 4,
                                      "resolution": [6, 7, 8, 5]}, [1, 2])
    with torch.no_grad():
        for planes in field.grids:
            for p in planes:
                p.copy_(torch.rand_like(p) + 0.25)
    g = torch.Generator().manual_seed(1)
    N = 40
    pts = torch.rand(N, 3, generator=g) * 2.4 - 1.2             # some outside [-1,1]: exercises the border clamp
    ts = torch.rand(N, 1, generator=g) * 2.2 - 1.1
    out = field(pts, ts)
    cot = torch.randn(out.shape, generator=g)
    grads = torch.autograd.grad((out * cot).sum(), [p for planes in field.grids for p in planes])
    blob = {"pts": pts, "ts": ts, "aabb": field.aabb.detach(), "out": out.detach(), "cot": cot}
    k = 0
    for s, planes in enumerate(field.grids):
        for p_i, p in enumerate(planes):
            blob[f"plane_{s}_{p_i}"] = p.detach()
            blob[f"grad_{s}_{p_i}"] = grads[k]
            k += 1
    np.savez_compressed(OUT / "hexplane.npz", **{k: v.numpy() for k, v in blob.items()})
    print("wrote hexplane.npz", tuple(out.shape))


if __name__ == "__main__":
    main()
