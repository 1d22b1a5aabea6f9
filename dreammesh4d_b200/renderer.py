"""Host-side mirror of the reference's renderer plugin for the dynamic stage.

``DiffGaussianBatchRenderer.batch_forward(batch)`` returns the same dict as
GaussianBatchRenderer.batch_forward + DiffGaussian.forward
(custom/threestudio-dreammesh4d/renderer/gaussian_batch_renderer.py:9-122,
renderer/diff_sugar_rasterizer_temporal.py:81-239) but renders every view of the batch in ONE launch
sequence: batched camera matrices, one fused skinning call for all timestamps, one 6-channel rasterizer
call (RGB + per-Gaussian normals share projection, binning and sort), no per-view host synchronisation
(no ``.item()``, no boolean-mask indexing) and one fused post-op kernel (``postops.py``) for the image tail.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import rasterizer as R
from .postops import post_ops
from .camera import get_cam_info_gaussian
from .geometry import DynamicSuGaRGeometry
from .nvtx import nvtx_range


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
        self._bg_cache = None

    def batch_forward(self, batch: Dict[str, Any], compute_normal_from_dist: bool = True, node_attrs=None) -> Dict[str, Any]:
        geo = self.geometry
        c2w = batch["c2w"]
        dev = c2w.device
        B = c2w.shape[0]
        H, W = int(batch["height"]), int(batch["width"])
        fovy = batch["fovy"]
        # gaussian_batch_renderer.py:23-26 — fovx := fovy, znear .1, zfar 100
        # "cam_info" lets a caller that replays this method from a CUDA graph hand in the camera block it computed
        # eagerly for the step (trainstep.GraphedDynamicStageStep); otherwise it is derived here
        view, proj, campos, tanx, tany = batch["cam_info"] if "cam_info" in batch else \
            get_cam_info_gaussian(c2w, fovy, fovy, znear=0.1, zfar=100.0)
        bgc = self.back_ground_color if self.training else tuple(1.0 - c for c in self.back_ground_color)   # temporal.py:96-103
        key = (bgc, str(dev))
        if self._bg_cache is None or self._bg_cache[0] != key:      # built once: no pageable H2D copy per step
            self._bg_cache = (key, torch.tensor(bgc + bgc, dtype=torch.float32, device=dev))
        bg6 = self._bg_cache[1].expand(B, 6)                                    # the normal pass uses the same bg

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
            with nvtx_range("dm4d.skin"):
                timed = geo.deform(batch["timestamp"], node_attrs=node_attrs)       # one set per view
            vp = R.make_view_params(view, proj, campos, tanx, tany, bg6, 1.0,
                                    set_index=torch.arange(B, device=dev))
            with nvtx_range("dm4d.rasterize"):
                color6, radii, depth, alpha = R.rasterize_batch(
                    timed["means3D"], geo.get_opacity, timed.get("scales", geo.get_scaling), timed["rotations"], geo.get_points_rgb(), vp, H, W,
                    colors2=timed["normals"], means2D=screenspace, capacity=self.capacity, distinct_sets=True,
                    state_out=states)
        self.last_state = states[0]
        self.last_view_params = vp        # [B,48] camera block actually rasterized (include/dm4d.h layout)
        # image post-ops (temporal.py:180-193,212-218,229 / normal.py:172-206) for the whole batch: one fused
        # forward kernel, outputs already [B,H,W,C]; its backward feeds the rasterizer backward directly
        nfd = compute_normal_from_dist and "rays_o" in batch
        with nvtx_range("dm4d.postops"):
            out = post_ops(color6, depth, alpha, batch["rays_o"] if nfd else None, batch["rays_d"] if nfd else None,
                           static=static, compute_normal_from_dist=nfd)
        out.update({
            "viewspace_points": screenspace,                                    # [B,P,3]; .grad holds the 2-D mean gradients
            "visibility_filter": list((radii > 0).unbind(0)),
            "radii": list(radii.unbind(0)),
        })
        return out
