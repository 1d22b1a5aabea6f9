"""Score-distillation step of the dynamic stage.

Host-side mirror of ``TemporalStableZero123Guidance`` (custom/threestudio-dreammesh4d/guidance/
temporal_stable_zero123_guidance.py): same method names, argument meaning and returned dict
(``__call__`` :299-374, ``get_cond`` :250-297, ``encode_images`` :228-236, ``set_min_max_steps`` :167-170,
``update_step`` :376-388), same order of random draws (posterior sample, timestep, noise), so a seeded CPU run
reproduces the reference's numbers (tests/test_sds.py executes the reference's own ``__call__`` source against a
stub network).  The network behind it is pluggable: anything with ``encode_first_stage`` /
``get_first_stage_encoding`` / ``cc_projection`` / ``apply_model`` — ``zero123.Zero123Model`` on the B200 path
(tensor-core matmuls; no hand-written kernels here by north-star).

What differs from the reference by design: no ``.item()`` / host-side branches inside the step (the clip value is a
Python float decided by ``update_step``, outside the step), random draws on the device, conditioning tensors cached
on the device once — the whole call is CUDA-graph capturable; and the camera-delta block ``T`` is built with one
stack instead of four separate small kernels.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Sequence, Union

import torch
import torch.nn.functional as F

from .nvtx import nvtx_range


def ddim_alphas_cumprod(num_train_timesteps: int, beta_start: float, beta_end: float) -> torch.Tensor:
    """``DDIMScheduler(..., beta_schedule="scaled_linear").alphas_cumprod`` as the reference builds it (:138-155):
    betas = linspace(sqrt(b0), sqrt(b1), T, fp32)^2 ; alphas_cumprod = cumprod(1 - betas).  diffusers is an
    un-vendored dependency; this is its published schedule."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(alphas_cumprod: torch.Tensor, x: torch.Tensor, noise: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """``DDIMScheduler.add_noise``: sqrt(ac[t]) x + sqrt(1 - ac[t]) noise (per-sample broadcast)."""
    ac = alphas_cumprod.to(device=x.device, dtype=x.dtype)[t]
    shape = (-1,) + (1,) * (x.dim() - 1)
    return ac.sqrt().reshape(shape) * x + (1.0 - ac).sqrt().reshape(shape) * noise


def scheduled(value: Union[float, Sequence[float]], epoch: int, global_step: int) -> float:
    """threestudio ``C(value, epoch, step)`` (threestudio/utils/misc.py:66-101): a number, or
    ``[start_step, v0, v1, end_step]`` linearly interpolated (integer end -> by step, float end -> by epoch)."""
    if isinstance(value, (int, float)):
        return float(value)
    v = list(value)
    if len(v) == 3:
        v = [0] + v
    if len(v) >= 6:                 # piecewise: [s0, v0, v1, s1, v2, s2, ...]
        sel = 3
        for i in range(3, len(v) - 2, 2):
            if global_step >= v[i]:
                sel = i + 2
        head = [v[sel - 2], v[sel - 3]] if sel != 3 else v[:2]
        v = head + [v[sel - 1], v[sel]]
    if len(v) != 4:
        raise ValueError("scheduled value must be a number or [start_step, start_value, end_value, end_step, ...]")
    start, v0, v1, end = v
    cur = global_step if isinstance(end, int) else epoch
    frac = max(min(1.0, (cur - start) / (end - start)), 0.0)
    return float(v0 + (v1 - v0) * frac)


class TemporalStableZero123SDS:
    """Drop-in for the registered ``temporal-stable-zero123-guidance`` object on the training path.

    ``model``: the latent-diffusion slice (``zero123.Zero123Model`` or any object with the four methods above).
    ``c_crossattn [n_frames,1,768]`` / ``c_concat [n_frames,4,32,32]``: the per-frame conditioning the reference caches
    in ``prepare_embeddings_video`` (:205-222) from the input video (CLIP image embedding, VAE mode of the frame).
    """

    def __init__(self, model, c_crossattn: torch.Tensor, c_concat: torch.Tensor, *, guidance_scale: float = 5.0,
                 cond_elevation_deg: float = 0.0, cond_azimuth_deg: float = 0.0, cond_camera_distance: float = 1.2,
                 min_step_percent: Union[float, Sequence[float]] = 0.02,
                 max_step_percent: Union[float, Sequence[float]] = 0.98, grad_clip: Optional[Any] = None,
                 weights_dtype: torch.dtype = torch.float16, num_train_timesteps: int = 1000,
                 linear_start: float = 0.00085, linear_end: float = 0.0120):
        self.model = model
        if isinstance(model, torch.nn.Module):
            for p in model.parameters():            # guidance :124-125 — the diffusion prior is frozen
                p.requires_grad_(False)
        self.weights_dtype = weights_dtype
        self.device = c_crossattn.device
        self.c_crossattn = c_crossattn.to(weights_dtype)
        self.c_concat = c_concat.to(weights_dtype)
        self.guidance_scale = float(guidance_scale)
        self.cond_elevation_deg = float(cond_elevation_deg)
        self.cond_azimuth_deg = float(cond_azimuth_deg)
        self.cond_camera_distance = float(cond_camera_distance)
        self.min_step_percent, self.max_step_percent, self.grad_clip = min_step_percent, max_step_percent, grad_clip
        self.grad_clip_val: Optional[float] = None
        self.num_train_timesteps = int(num_train_timesteps)
        self.alphas = ddim_alphas_cumprod(num_train_timesteps, linear_start, linear_end).to(self.device)
        self.set_min_max_steps()

    # ---- schedule (:167-170, :376-388) ---------------------------------------------------------------------------
    def set_min_max_steps(self, min_step_percent: float = 0.02, max_step_percent: float = 0.98) -> None:
        self.min_step = int(self.num_train_timesteps * min_step_percent)
        self.max_step = int(self.num_train_timesteps * max_step_percent)

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False) -> None:
        if self.grad_clip is not None:
            self.grad_clip_val = scheduled(self.grad_clip, epoch, global_step)
        self.set_min_max_steps(min_step_percent=scheduled(self.min_step_percent, epoch, global_step),
                               max_step_percent=scheduled(self.max_step_percent, epoch, global_step))

    # ---- pieces of the step --------------------------------------------------------------------------------------
    def encode_images(self, imgs: torch.Tensor, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """:228-236 — [B,3,256,256] in [0,1] -> latents [B,4,32,32] (gradient flows to ``imgs``)."""
        x = (imgs * 2.0 - 1.0).to(self.weights_dtype)
        post = self.model.encode_first_stage(x)
        return self.model.get_first_stage_encoding(post, generator).to(imgs.dtype)

    @torch.no_grad()
    def get_cond(self, elevation: torch.Tensor, azimuth: torch.Tensor, camera_distances: torch.Tensor,
                 frame_indices: torch.Tensor, c_crossattn: Optional[torch.Tensor] = None,
                 c_concat: Optional[torch.Tensor] = None, **kwargs) -> Dict[str, list]:
        """:250-297 — camera delta [polar difference, sin/cos of the azimuth difference, conditioning polar angle]
        appended to the frame's CLIP embedding and projected; unconditional half = zeros, batch order [uncond | cond]."""
        d_az = torch.deg2rad(azimuth - self.cond_azimuth_deg)
        T = torch.stack([torch.deg2rad((90.0 - elevation) - (90.0 - self.cond_elevation_deg)),
                         torch.sin(d_az), torch.cos(d_az),
                         torch.deg2rad(90.0 - torch.full_like(elevation, self.cond_elevation_deg))], dim=-1)
        T = T[:, None, :].to(device=self.device, dtype=self.weights_dtype)
        cc = self.c_crossattn if c_crossattn is None else c_crossattn
        cat = self.c_concat if c_concat is None else c_concat
        clip_emb = self.model.cc_projection(torch.cat([cc[frame_indices], T], dim=-1))
        cond_cat = cat[frame_indices]
        return {"c_crossattn": [torch.cat([torch.zeros_like(clip_emb), clip_emb], dim=0)],
                "c_concat": [torch.cat([torch.zeros_like(cond_cat), cond_cat], dim=0)]}

    # ---- the step (:299-374) -------------------------------------------------------------------------------------
    def __call__(self, rgb: torch.Tensor, elevation: torch.Tensor, azimuth: torch.Tensor,
                 camera_distances: torch.Tensor, frame_indices: torch.Tensor, rgb_as_latents: bool = False,
                 generator: Optional[torch.Generator] = None, **kwargs) -> Dict[str, Any]:
        B = rgb.shape[0]
        rgb_bchw = rgb.permute(0, 3, 1, 2)
        if rgb_as_latents:
            latents = F.interpolate(rgb_bchw, (32, 32), mode="bilinear", align_corners=False) * 2 - 1
        else:
            with nvtx_range("dm4d.sds.encode"):
                latents = self.encode_images(F.interpolate(rgb_bchw, (256, 256), mode="bilinear", align_corners=False),
                                             generator)
        cond = self.get_cond(elevation, azimuth, camera_distances, frame_indices)
        t = torch.randint(self.min_step, self.max_step + 1, [B], dtype=torch.long, device=latents.device,
                          generator=generator)
        with torch.no_grad():
            noise = torch.empty_like(latents).normal_(generator=generator)      # == randn_like (follows latents' memory format)
            noisy = add_noise(self.alphas, latents, noise, t)
            with nvtx_range("dm4d.sds.unet"):
                eps = self.model.apply_model(torch.cat([noisy, noisy]).to(self.weights_dtype), torch.cat([t, t]), cond)
            eps_uncond, eps_cond = eps.chunk(2)
            eps = eps_uncond + self.guidance_scale * (eps_cond - eps_uncond)
            w = (1.0 - self.alphas.to(latents.device)[t]).reshape(-1, 1, 1, 1)
            grad = torch.nan_to_num(w * (eps - noise))
            if self.grad_clip_val is not None:
                grad = grad.clamp(-self.grad_clip_val, self.grad_clip_val)
            target = latents - grad
        # d loss / d latents == grad (the reference's reparameterisation of SpecifyGradient, :363-365)
        loss_sds = 0.5 * F.mse_loss(latents, target, reduction="sum") / B
        return {"loss_sds": loss_sds, "grad_norm": grad.norm(), "min_step": self.min_step, "max_step": self.max_step}
