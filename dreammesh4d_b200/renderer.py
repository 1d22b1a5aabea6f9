"""Host-side mirror of the reference's renderer plugin for the dynamic stage.

``DiffGaussianBatchRenderer.batch_forward(batch)`` returns the same dict as
GaussianBatchRenderer.batch_forward + DiffGaussian.forward
(custom/threestudio-dreammesh4d/renderer/gaussian_batch_renderer.py:9-122,
renderer/diff_sugar_rasterizer_temporal.py:81-239) but renders every view of the batch in ONE launch
sequence: batched camera matrices, one fused skinning call for all timestamps, one 6-channel rasterizer
call (RGB + per-Gaussian normals share projection, binning and sort), no per-view host synchronisation
(no ``.item()``, no boolean-mask indexing).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F

from . import rasterizer as R
from .camera import get_cam_info_gaussian
from .geometry import DynamicSuGaRGeometry


def depth_to_normal(xyz_map: torch.Tensor) -> torch.Tensor:
    """Depth2Normal (diff_sugar_rasterizer_temporal.py:25-54): central differences with zero padding on
    a [B,3,H,W] position map, normal = -cross(d/dx, d/dy)."""
    p = F.pad(xyz_map, (1, 1, 1, 1))
    ddx = p[:, :, 1:-1, 2:] - p[:, :, 1:-1, :-2]
    ddy = p[:, :, 2:, 1:-1] - p[:, :, :-2, 1:-1]
    return -torch.cross(ddx, ddy, dim=1)


def _detach_outside(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """x[~mask] = x[~mask].detach() without boolean indexing (no host sync)."""
    return torch.where(mask, x, x.detach())


class DiffGaussianBatchRenderer:
    """Mirror of the registered ``diff-sugar-rasterizer-temporal`` renderer (temporal.py:56-79)."""

    def __init__(self, geometry: DynamicSuGaRGeometry, back_ground_color=(1.0, 1.0, 1.0), training: bool = True,
                 capacity: Optional[int] = None):
        self.geometry = geometry
        self.back_ground_color = tuple(float(c) for c in back_ground_color)
        self.training = training
        self.capacity = capacity          # None: exact sizing with one num_rendered read-back per batch
        self.last_state = None
        self.last_view_params = None

    def batch_forward(self, batch: Dict[str, Any], compute_normal_from_dist: bool = True, node_attrs=None) -> Dict[str, Any]:
        geo = self.geometry
        c2w = batch["c2w"]
        dev = c2w.device
        B = c2w.shape[0]
        H, W = int(batch["height"]), int(batch["width"])
        fovy = batch["fovy"]
        # gaussian_batch_renderer.py:23-26 — fovx := fovy, znear .1, zfar 100
        view, proj, campos, tanx, tany = get_cam_info_gaussian(c2w, fovy, fovy, znear=0.1, zfar=100.0)
        bg = torch.tensor(self.back_ground_color, dtype=torch.float32, device=dev)
        if not self.training:
            bg = 1.0 - bg                                                       # temporal.py:96-103
        bg6 = torch.cat([bg, bg]).expand(B, 6)                                  # the normal pass uses the same bg

        static = "timestamp" not in batch       # static stage: diff_sugar_rasterizer_normal.py:79-226
        P = geo.n_gaussians
        screenspace = torch.zeros(B, P, 3, dtype=torch.float32, device=dev, requires_grad=True)   # temporal.py:108-113
        states = []
        if static:
            # one shared attribute set for every view (SuGaRModel getters, sugar.py:528-546)
            vp = R.make_view_params(view, proj, campos, tanx, tany, bg6, 1.0)
            colors = batch.get("override_color", None)
            if colors is None:
                colors = geo.get_points_rgb()                                   # sugar_static.py:90-94 injects exactly this
            color6, radii, depth, alpha = R.rasterize_batch(
                geo.get_xyz, geo.get_opacity, geo.get_scaling, geo.get_rotation, colors, vp, H, W,
                colors2=geo.get_gs_normals, means2D=screenspace, capacity=self.capacity, state_out=states)
        else:
            timed = geo.deform(batch["timestamp"], node_attrs=node_attrs)       # one set per view
            vp = R.make_view_params(view, proj, campos, tanx, tany, bg6, 1.0,
                                    set_index=torch.arange(B, device=dev))
            color6, radii, depth, alpha = R.rasterize_batch(
                timed["means3D"], geo.get_opacity, geo.get_scaling, timed["rotations"], geo.get_points_rgb(), vp, H, W,
                colors2=timed["normals"], means2D=screenspace, capacity=self.capacity, distinct_sets=True,
                state_out=states)
        self.last_state = states[0]
        self.last_view_params = vp        # [B,48] camera block actually rasterized (include/dm4d.h layout)
        rgb, nrm = color6[:, :3], color6[:, 3:]

        mask = alpha > 0.99                                                     # temporal.py:180
        depth = _detach_outside(depth, mask)                                    # :181
        mask3 = mask.expand(-1, 3, -1, -1)
        out = {}
        if compute_normal_from_dist:
            rays_o = batch["rays_o"].permute(0, 3, 1, 2)
            rays_d = batch["rays_d"].permute(0, 3, 1, 2)
            xyz_map = rays_o + depth * rays_d                                   # :187
            nfd = F.normalize(depth_to_normal(xyz_map), dim=1)                  # :188-189
            nfd_map = nfd * 0.5 * alpha + 0.5                                   # :190
            out["comp_normal_from_dist"] = _detach_outside(nfd_map, mask3).permute(0, 2, 3, 1)   # :191-193
        normal = F.normalize(nrm, dim=1)                                        # :212
        normal_map = normal * 0.5 * alpha + 0.5                                 # :214
        out.update({
            "comp_rgb": rgb.clamp(0, 1).permute(0, 2, 3, 1),                    # :229, batch renderer :79
            "comp_normal": _detach_outside(normal_map, mask3).permute(0, 2, 3, 1),
            "comp_depth": depth.permute(0, 2, 3, 1),
            "comp_mask": alpha.permute(0, 2, 3, 1),
            "viewspace_points": screenspace,                                    # [B,P,3]; .grad holds the 2-D mean gradients
            "visibility_filter": [radii[b] > 0 for b in range(B)],
            "radii": [radii[b] for b in range(B)],
        })
        return out
