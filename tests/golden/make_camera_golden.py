#!/usr/bin/env python
"""Generates tests/golden/camera.npz by executing the reference's own camera functions
(threestudio/utils/ops.py: convert_pose, get_projection_matrix_gaussian, get_cam_info_gaussian — unmodified source text,
extracted by AST from /root/reference at generation time only) on seeded orbit cameras.  The functions hard-code
device="cuda" / .cuda(); they are run on the CPU by handing them a `torch` proxy whose zeros/eye ignore the device
argument and by making Tensor.cuda the identity during the call — the arithmetic is untouched."""
import ast
import math
import sys
import types
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent
ROOT = OUT.parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/threestudio/utils/ops.py")
WANT = {"convert_pose", "get_projection_matrix_gaussian", "get_cam_info_gaussian"}


class _TorchProxy(types.ModuleType):
    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def zeros(*a, device=None, **k):
        return torch.zeros(*a, **k)

    @staticmethod
    def eye(*a, device=None, **k):
        return torch.eye(*a, **k)


def main():
    src = REF.read_text()
    ns = {"torch": _TorchProxy("torch_proxy"), "math": math}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in WANT:
            exec(compile(ast.Module([node], []), str(REF), "exec"), ns)
    from dreammesh4d_b200 import synthetic
    B = 6
    c2w, fovy = synthetic.random_orbit_cameras(B, seed=21, fovy_deg=20.0)
    fovy = fovy * torch.linspace(0.8, 1.6, B)                     # different fields of view
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        outs = [ns["get_cam_info_gaussian"](c2w=c2w[b], fovx=float(fovy[b]), fovy=float(fovy[b]), znear=0.1, zfar=100.0)
                for b in range(B)]
    finally:
        torch.Tensor.cuda = saved
    np.savez_compressed(OUT / "camera.npz", c2w=c2w.numpy(), fovy=fovy.numpy(),
                        world_view=torch.stack([o[0] for o in outs]).numpy(), full_proj=torch.stack([o[1] for o in outs]).numpy(),
                        center=torch.stack([o[2] for o in outs]).numpy())
    print("wrote camera.npz")


if __name__ == "__main__":
    main()
