#!/usr/bin/env python
"""Generates tests/golden/sds.npz by executing the reference's OWN score-distillation code (unmodified source text of
``TemporalStableZero123Guidance.__call__`` / ``get_cond`` / ``encode_images`` / ``set_min_max_steps``, read from
/root/reference/custom/threestudio-dreammesh4d/guidance/temporal_stable_zero123_guidance.py at generation time only;
the class itself cannot be imported here — diffusers, omegaconf, clip, threestudio are absent).

The methods are bound to a stand-in object that provides exactly the attributes they read:
  * ``self.model``: a small latent-diffusion stub with the four entry points the methods call — its
    ``encode_first_stage`` returns the reference's own ``DiagonalGaussianDistribution`` (imported from
    /root/reference/extern/ldm_zero123/modules/distributions/distributions.py) so the posterior sampling is the
    reference's; the networks are seeded random convolutions (what they compute is irrelevant to the SDS arithmetic).
  * ``self.scheduler.add_noise`` / ``self.alphas``: diffusers' DDIMScheduler is un-vendored; restated from its
    published formulas (scaled_linear betas, sqrt(ac) x + sqrt(1-ac) eps) — the one UNPINNED third-party piece.
Outputs: loss_sds, grad_norm, d loss / d rgb, the conditioning tensors, for two configurations (with / without
grad clipping, different guidance scales and step ranges).
"""
import ast
import sys
import textwrap
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

OUT = Path(__file__).resolve().parent
REF = Path("/root/reference/custom/threestudio-dreammesh4d/guidance/temporal_stable_zero123_guidance.py")
sys.path.insert(0, str(OUT.parents[1]))
from tests.helpers import SDSStubModel, sds_stub_inputs  # noqa: E402


def reference_methods():
    src = REF.read_text()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TemporalStableZero123Guidance")
    lines = src.splitlines()
    chunks = []
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("__call__", "get_cond", "encode_images", "set_min_max_steps"):
            first = min([fn.lineno] + [d.lineno for d in fn.decorator_list])
            chunks.append(textwrap.dedent("\n".join(lines[first - 1:fn.end_lineno])))
    ns = {"torch": torch, "F": F}
    exec("from __future__ import annotations\n" + "\n\n".join(chunks), ns)      # annotations stay unevaluated strings
    return {k: ns[k] for k in ("__call__", "get_cond", "encode_images", "set_min_max_steps")}


def reference_posterior_class():
    p = Path("/root/reference/extern/ldm_zero123/modules/distributions/distributions.py")
    ns = {}
    exec(compile(p.read_text(), str(p), "exec"), ns)
    return ns["DiagonalGaussianDistribution"]


class DDIMStub:
    """diffusers.DDIMScheduler as configured at guidance :138-146 (restated; un-vendored)."""

    def __init__(self, n, b0, b1):
        betas = torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        a = ac[timesteps] ** 0.5
        s = (1 - ac[timesteps]) ** 0.5
        while len(a.shape) < len(original_samples.shape):
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise


def run(methods, DGD, case):
    model = SDSStubModel(seed=3)
    post_cls = DGD

    class RefModel:            # the stub network behind the reference's calling convention
        cc_projection = model.cc_projection
        scale_factor = model.scale_factor

        def encode_first_stage(self, x):
            return post_cls(model.moments(x))

        def get_first_stage_encoding(self, post):
            return self.scale_factor * post.sample()

        def apply_model(self, x, t, cond):
            return model.apply_model(x, t, cond)

    inp = sds_stub_inputs(seed=case["seed"])
    self = types.SimpleNamespace()
    self.cfg = types.SimpleNamespace(cond_elevation_deg=case["cond_elev"], cond_azimuth_deg=case["cond_azim"],
                                     cond_camera_distance=1.2, guidance_scale=case["scale"])
    self.model, self.device, self.weights_dtype = RefModel(), torch.device("cpu"), torch.float32
    self.scheduler = DDIMStub(1000, 0.00085, 0.0120)
    self.alphas = self.scheduler.alphas_cumprod
    self.num_train_timesteps = 1000
    self.grad_clip_val = case["clip"]
    self.c_crossattn, self.c_concat = inp["c_crossattn"], inp["c_concat"]
    for name in ("get_cond", "encode_images"):
        setattr(self, name, types.MethodType(methods[name], self))
    methods["set_min_max_steps"](self, case["min_pct"], case["max_pct"])
    rgb = inp["rgb"].clone().requires_grad_(True)
    torch.manual_seed(case["seed"] + 100)
    out = methods["__call__"](self, rgb, inp["elevation"], inp["azimuth"], inp["camera_distances"], inp["frame_indices"])
    out["loss_sds"].backward()
    cond = self.get_cond(inp["elevation"], inp["azimuth"], inp["camera_distances"], inp["frame_indices"])
    return {"loss_sds": out["loss_sds"].detach(), "grad_norm": out["grad_norm"].detach(), "d_rgb": rgb.grad,
            "min_step": torch.tensor(out["min_step"]), "max_step": torch.tensor(out["max_step"]),
            "cond_crossattn": cond["c_crossattn"][0], "cond_concat": cond["c_concat"][0]}


CASES = [dict(seed=1, scale=3.0, clip=None, min_pct=0.02, max_pct=0.5, cond_elev=0.0, cond_azim=0.0),
         dict(seed=2, scale=5.0, clip=0.05, min_pct=0.02, max_pct=0.98, cond_elev=10.0, cond_azim=-30.0)]


def main():
    methods, DGD = reference_methods(), reference_posterior_class()
    blob = {}
    for i, case in enumerate(CASES):
        for k, v in run(methods, DGD, case).items():
            blob[f"c{i}_{k}"] = v.numpy()
        print(f"case {i}: loss_sds {float(blob[f'c{i}_loss_sds']):.6f} grad_norm {float(blob[f'c{i}_grad_norm']):.6f} "
              f"|d_rgb|max {np.abs(blob[f'c{i}_d_rgb']).max():.3e}")
    np.savez_compressed(OUT / "sds.npz", **blob)


if __name__ == "__main__":
    main()
