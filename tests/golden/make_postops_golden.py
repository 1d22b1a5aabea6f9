#!/usr/bin/env python
"""Generates tests/golden/postops.npz by executing the reference's OWN post-op statements (unmodified source text,
read from /root/reference at generation time only):

  * class Depth2Normal of custom/threestudio-dreammesh4d/renderer/diff_sugar_rasterizer_temporal.py (:25-54), and
  * the statements of DiffGaussian.forward between the first rasterizer call and the return dict of
    diff_sugar_rasterizer_temporal.py (:180-218) and diff_sugar_rasterizer_normal.py (:172-206),

with the two ``rasterizer(...)`` calls replaced by a stub that returns seeded images (the rasterizer itself is the
un-vendored CUDA module).  Outputs and autograd gradients w.r.t. the four rasterizer images are committed.
"""
import ast
import sys
import textwrap
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

OUT = Path(__file__).resolve().parent
REFDIR = Path("/root/reference/custom/threestudio-dreammesh4d/renderer")
H, W = 20, 24


def extract(path: Path):
    src = path.read_text()
    tree = ast.parse(src)
    d2n = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Depth2Normal")
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("DiffGaussian", "DiffSuGaR"))
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "forward")
    # statements from the first `rendered_image, radii, rendered_depth, rendered_alpha = rasterizer(...)` (inclusive)
    # up to the final `return {...}` (inclusive)
    start = next(i for i, st in enumerate(fwd.body) if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Tuple)
                 and [getattr(e, "id", None) for e in st.targets[0].elts][:1] == ["rendered_image"])
    lines = src.splitlines()
    body = "\n".join(lines[fwd.body[start].lineno - 1:fwd.body[-1].end_lineno])     # verbatim source lines
    return ast.get_source_segment(src, d2n), textwrap.dedent(body)


def run(path: Path, static: bool, seed: int):
    d2n_src, body = extract(path)
    ns = {"torch": torch, "F": F}
    exec(d2n_src, ns)
    g = torch.Generator().manual_seed(seed)
    rgb = (torch.rand(3, H, W, generator=g, dtype=torch.float64) * 1.4 - 0.2).requires_grad_(True)   # some outside [0,1]
    nrm = torch.randn(3, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    depth = (3.3 + 0.3 * torch.rand(1, H, W, generator=g, dtype=torch.float64)).requires_grad_(True)
    alpha = torch.rand(1, H, W, generator=g, dtype=torch.float64)
    alpha = torch.where(alpha > 0.35, 0.99 + 0.01 * alpha, alpha).requires_grad_(True)               # ~65 % inside the mask
    rays_o = torch.randn(3, generator=g, dtype=torch.float64).expand(H, W, 3).contiguous()
    rays_d = F.normalize(torch.randn(H, W, 3, generator=g, dtype=torch.float64) * 0.1 + torch.tensor([0.0, 0.0, -1.0]), dim=-1)
    calls = []

    def rasterizer(**kw):
        calls.append(kw)
        if len(calls) == 1:
            return rgb * 1.0, torch.ones(7, dtype=torch.int32), depth * 1.0, alpha * 1.0
        return nrm * 1.0, None, None, None

    class _Self:
        training = False
        normal_module = ns["Depth2Normal"]()
        # Depth2Normal builds fp32 kernels; run the reference statements in fp64 by promoting them (values are 0/±1)
    _Self.normal_module.delzdelxkernel = _Self.normal_module.delzdelxkernel.double()
    _Self.normal_module.delzdelykernel = _Self.normal_module.delzdelykernel.double()

    class _PC:
        get_gs_normals = torch.zeros(7, 3)

        @staticmethod
        def get_timed_gs_normals(*a, **k):
            return torch.zeros(1, 7, 3)

    class _Cam:
        timestamp = torch.tensor(0.5)
        frame_idx = torch.tensor(0)

    means2D = torch.zeros(7, 3)
    env = dict(ns, rasterizer=rasterizer, self=_Self(), pc=_PC(), viewpoint_camera=_Cam(), static=static,
               compute_normal_from_dist=True, kwargs={"batch_idx": 0, "rays_d": rays_d[None], "rays_o": rays_o[None]},
               means3D=None, means2D=means2D, shs=None, colors_precomp=None, opacity=None, scales=None, rotations=None,
               cov3D_precomp=None, screenspace_points=means2D)
    code = "def _f():\n" + "\n".join("    " + l for l in body.splitlines()) + "\n_out = _f()\n"
    exec(code, env)
    out = env["_out"]
    assert len(calls) == 2
    keys = ["render", "normal", "normal_from_dist", "depth", "mask"]
    cot = {k: torch.randn(out[k].shape, generator=g, dtype=torch.float64) for k in keys}
    loss = sum((out[k] * cot[k]).sum() for k in keys)
    grads = torch.autograd.grad(loss, (rgb, nrm, depth, alpha))
    rec = {"rgb": rgb, "nrm": nrm, "depth": depth, "alpha": alpha, "rays_o": rays_o, "rays_d": rays_d}
    rec.update({f"out_{k}": out[k] for k in keys})
    rec.update({f"cot_{k}": cot[k] for k in keys})
    rec.update({f"grad_{k}": gr for k, gr in zip(("rgb", "nrm", "depth", "alpha"), grads)})
    return {k: v.detach().numpy() for k, v in rec.items()}


def main():
    blob = {}
    for tag, fname, static in (("temporal", "diff_sugar_rasterizer_temporal.py", False),
                               ("static", "diff_sugar_rasterizer_normal.py", True)):
        for k, v in run(REFDIR / fname, static, seed=3 if static else 2).items():
            blob[f"{tag}_{k}"] = v
    np.savez_compressed(OUT / "postops.npz", **blob)
    print("wrote postops.npz", sorted(blob)[:6], "...")


if __name__ == "__main__":
    main()
