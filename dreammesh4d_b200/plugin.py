"""B-outer drop-in: the four classes DreamMesh4D registers with threestudio, backed by the B200 hot path.

    "sugar"                            <- custom/threestudio-dreammesh4d/geometry/sugar.py:33            (SuGaRModel)
    "dynamic-sugar"                    <- .../geometry/dynamic_sugar.py:42                               (DynamicSuGaRModel)
    "diff-sugar-rasterizer-normal"     <- .../renderer/diff_sugar_rasterizer_normal.py:54                (DiffSuGaR)
    "diff-sugar-rasterizer-temporal"   <- .../renderer/diff_sugar_rasterizer_temporal.py:56              (DiffGaussian)
(The guidance object "temporal-stable-zero123-guidance" keeps its reference registration: it needs the Zero123 checkpoint and
the CLIP image encoder at configure time; ``sds.TemporalStableZero123SDS`` is its per-step mirror and takes over ``.model``.)

Same construction protocol (``cls(cfg_dict, geometry=..., material=..., background=...)`` -> ``configure``,
threestudio/utils/base.py:96-115, threestudio/systems/base.py:292-303), the same ``Config`` fields and defaults (so the
shipped YAMLs parse unchanged; unknown keys raise as OmegaConf's structured configs do), the reference's parameter /
buffer names (its checkpoints load), and the calls the systems make: ``renderer.batch_forward(batch)``,
``geometry.update_learning_rate / optimizer / merge_optimizer / update_step / get_xyz / get_xyz_verts / get_faces /
get_timed_surface_mesh / get_timed_vertex_xyz / get_timed_vertex_rotation / _deformed_vert_positions``
(system/sugar_4dgen.py:66-81, 214-225, 286-297, 372-395; system/sugar_static.py:90-94).

With threestudio importable the classes derive from its ``BaseGeometry`` / ``Rasterizer`` and ``install()`` puts them in
its registry, replacing (or pre-empting) the reference's own registrations — ``launch.py``, the plugin and the YAMLs
stay untouched; see INTEGRATION.md.  Without threestudio (this repository's tests) a minimal stand-in base with the
same constructor protocol is used, so the classes are exercised exactly as the systems would.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import mesh_io
from .deform_graph import build_deformation_graph, sample_surface_points
from .deformation import HexPlaneDeformation
from .geometry import SuGaRState
from .renderer import DiffGaussianBatchRenderer

REGISTRY: Dict[str, type] = {}          # what this module provides, by registered name

try:                                     # the real host framework
    import threestudio                                         # noqa: F401
    from threestudio.models.geometry.base import BaseGeometry as _TSGeometry
    from threestudio.models.renderers.base import Rasterizer as _TSRasterizer
    HAVE_THREESTUDIO = True
except Exception:                        # stand-in with the same constructor protocol (threestudio/utils/base.py:96-115)
    HAVE_THREESTUDIO = False

    def _parse_structured(config_cls, cfg: Optional[dict]):
        cfg = dict(cfg or {})
        names = {f.name for f in dataclasses.fields(config_cls)}
        unknown = set(cfg) - names
        if unknown:
            raise KeyError(f"{config_cls.__qualname__}: unknown config key(s) {sorted(unknown)}")
        return config_cls(**cfg)

    class _StandInModule(nn.Module):
        @dataclass
        class Config:
            weights: Optional[str] = None

        def __init__(self, cfg: Optional[dict] = None, *args, **kwargs) -> None:
            super().__init__()
            self.cfg = _parse_structured(self.Config, cfg)
            self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
            self.configure(*args, **kwargs)

        def configure(self, *args, **kwargs) -> None:
            pass

        def do_update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
            self.update_step(epoch, global_step, on_load_weights)

        def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False):
            pass

    class _TSGeometry(_StandInModule):
        @dataclass
        class Config(_StandInModule.Config):
            pass

    class _TSRasterizer(_StandInModule):
        @dataclass
        class Config(_StandInModule.Config):
            radius: float = 1.0

        def configure(self, geometry, material=None, background=None) -> None:
            @dataclass
            class SubModules:                 # non-owning references, like threestudio/models/renderers/base.py:28-35
                geometry: Any
                material: Any
                background: Any
            self.sub_modules = SubModules(geometry, material, background)

        @property
        def geometry(self):
            return self.sub_modules.geometry


def _provides(name: str):
    def deco(cls):
        REGISTRY[name] = cls
        return cls
    return deco


# ------------------------------------------------------------------------------------------------------------------
# geometry
# ------------------------------------------------------------------------------------------------------------------
@_provides("sugar")
class SuGaRModel(SuGaRState, _TSGeometry):
    """Static surface-bound Gaussians (stage 2 of the pipeline, configs/sugar_static_refine.yaml)."""

    @dataclass
    class Config(_TSGeometry.Config):          # geometry/sugar.py:35-69, same names and defaults
        sh_levels: int = 1
        position_lr: Any = 0.001
        feature_lr: Any = 0.01
        opacity_lr: Any = 0.05
        scaling_lr: Any = 0.005
        rotation_lr: Any = 0.005
        learnable_positions: bool = False
        triangle_scale: float = 1.0
        n_gaussians_per_surface_triangle: int = 1
        keep_track_of_knn: bool = False
        knn_to_track: int = 16
        beta_mode: str = "average"
        primitive_types: str = "diamond"
        surface_mesh_to_bind_path: str = ""
        learn_surface_mesh_positions: bool = True
        learn_surface_mesh_opacity: bool = True
        learn_surface_mesh_scales: bool = True
        freeze_gaussians: bool = False
        spatial_lr_scale: float = 10.0
        spatial_extent: float = 3.5
        color_clip: Any = 2.0
        gs_color_inherit_vertices: bool = True
        init_gs_opacity: float = 0.5
        geometry_convert_from: str = ""
        square_size_in_texture: int = 10
        pred_normal: bool = False
        init_gs_scales_s: float = 1.7

    cfg: Config

    def _bind_mesh(self, o3d_mesh=None) -> mesh_io.SugarScene:
        c = self.cfg
        source = o3d_mesh if o3d_mesh is not None else c.surface_mesh_to_bind_path
        if isinstance(source, mesh_io.SugarScene):
            return source
        if isinstance(source, str) and not source:
            raise ValueError("surface_mesh_to_bind_path is empty and no mesh object was passed to configure()")
        return mesh_io.load_scene(source, c.n_gaussians_per_surface_triangle, init_gs_scales_s=c.init_gs_scales_s,
                                  init_gs_opacity=c.init_gs_opacity, spatial_extent=c.spatial_extent,
                                  inherit_vertex_colors=c.gs_color_inherit_vertices, learn_opacity=c.learn_surface_mesh_opacity)

    def _learn_flags(self) -> Dict[str, bool]:
        c = self.cfg                            # sugar.py:168-171, 227-233
        return dict(points=c.learn_surface_mesh_positions, scales=c.learn_surface_mesh_scales,
                    quaternions=c.learn_surface_mesh_scales, densities=c.learn_surface_mesh_opacity, sh=not c.freeze_gaussians)

    def configure(self, o3d_mesh=None) -> None:
        super().configure()
        self.active_sh_degree, self.sh_levels = 0, self.cfg.sh_levels
        self._install_state(self._bind_mesh(o3d_mesh), None, None, sh_levels=self.cfg.sh_levels, learn=self._learn_flags())
        self.to(self.device)
        self.spatial_lr_scale = self.cfg.spatial_lr_scale
        self._set_optimizer(self._static_groups(self.cfg))
        self.color_clip = float(self.cfg.color_clip) if isinstance(self.cfg.color_clip, (int, float)) else None

    save_path = None                    # debug attribute the systems assign (sugar_4dgen.py)
    pruned_or_densified = False         # free-Gaussian densification (sugar_static.py:109) does not exist for bound Gaussians

    def training_setup(self) -> None:
        """sugar.py:327-385 — (re)build the optimizer groups."""
        self._set_optimizer(self._static_groups(self.cfg))

    def update_step(self, epoch: int = 0, global_step: int = 0, on_load_weights: bool = False):
        SuGaRState.update_step(self, epoch, global_step, on_load_weights)


@_provides("dynamic-sugar")
class DynamicSuGaRModel(SuGaRModel):
    """Mesh + Gaussians driven by a sparse control graph and the HexPlane deformation network (stage 3,
    configs/sugar_dynamic_dg.yaml)."""

    @dataclass
    class Config(SuGaRModel.Config):           # geometry/dynamic_sugar.py:44-73
        num_frames: int = 14
        static_learnable: bool = False
        use_deform_graph: bool = True
        dynamic_mode: str = "deformation"
        n_dg_nodes: int = 1000
        dg_node_connectivity: int = 8
        dg_trans_lr: Any = 0.001
        dg_rot_lr: Any = 0.001
        dg_scale_lr: Any = 0.001
        vert_trans_lr: Any = 0.001
        vert_rot_lr: Any = 0.001
        vert_scale_lr: Any = 0.001
        deformation_lr: Any = 0.001
        grid_lr: Any = 0.001
        d_xyz: bool = True
        d_rotation: bool = True
        d_opacity: bool = False
        d_scale: bool = True
        dist_mode: str = "eucdisc"
        skinning_method: str = "hybrid"

    cfg: Config

    def configure(self, o3d_mesh=None, xyz_nodes: Optional[torch.Tensor] = None) -> None:
        c = self.cfg
        if c.dynamic_mode != "deformation" or not c.use_deform_graph:
            # 'discrete' + hybrid raises in the reference itself (typo at dynamic_sugar.py:118-120); per-vertex mode is
            # a different parametrisation (no control graph) that the hot path does not cover
            raise NotImplementedError("dreammesh4d_b200 covers dynamic_mode='deformation' with use_deform_graph=True "
                                      "(the configuration of configs/sugar_dynamic_dg.yaml)")
        if c.d_scale and c.skinning_method == "dqs":
            raise ValueError("d_scale=True with skinning_method='dqs': the reference produces no vertex scale there "
                             "(dynamic_sugar.py:595-612) and fails with a KeyError at :699")
        _TSGeometry.configure(self)
        self.active_sh_degree, self.sh_levels = 0, c.sh_levels
        scene = self._bind_mesh(o3d_mesh)
        learn = self._learn_flags() if c.static_learnable else {k: False for k in ("points", "scales", "quaternions", "densities", "sh")}
        # dynamic_sugar.py:141-148: scale head unless (dqs and not d_scale), opacity (LBS weight) head only for hybrid
        net = HexPlaneDeformation(no_ds=not (c.d_scale or c.skinning_method in ("hybrid", "lbs")),
                                  no_do=c.skinning_method != "hybrid")
        self._install_state(scene, None, net, skinning_method=c.skinning_method, static_learnable=c.static_learnable,
                            sh_levels=c.sh_levels, learn=learn)
        self.to(self.device)
        self.d_scale = bool(c.d_scale)            # per-Gaussian scale deformation: tensor ops on top of the fused kernels
        self.num_frames, self.dynamic_mode = c.num_frames, c.dynamic_mode
        self.build_deformation_graph(c.n_dg_nodes, xyz_nodes, nodes_connectivity=c.dg_node_connectivity, mode=c.dist_mode)
        self.spatial_lr_scale = c.spatial_lr_scale
        self._set_optimizer((self._static_groups(c) if c.static_learnable else []) + self._dynamic_groups(c))
        self.color_clip = float(c.color_clip) if isinstance(c.color_clip, (int, float)) else None

    def build_deformation_graph(self, n_nodes: int, xyz_nodes=None, nodes_connectivity: int = 6, mode: str = "geodisc",
                                seed: int = 0) -> None:
        """dynamic_sugar.py:745-861 on the GPU (deform_graph.py): area-uniform node sampling unless ``xyz_nodes`` is given,
        K nearest nodes per vertex (Euclidean or geodesic), the reference's weight formulas."""
        verts, faces = self._points.detach(), self._surface_mesh_faces
        nodes = sample_surface_points(verts, faces, n_nodes, seed=seed) if xyz_nodes is None else xyz_nodes.to(verts)
        graph, conn = build_deformation_graph(verts, nodes, nodes_connectivity, mode=mode, faces=faces)
        self._install_graph(graph)
        self._deform_graph_connectivity = conn


# ------------------------------------------------------------------------------------------------------------------
# renderers
# ------------------------------------------------------------------------------------------------------------------
class _BatchedRasterizer(_TSRasterizer):
    """Both registered renderers: one batched 6-channel pass + fused post-ops for all views of the batch
    (``renderer.DiffGaussianBatchRenderer``) behind ``batch_forward``; ``forward`` renders a single view the same way."""

    @dataclass
    class Config(_TSRasterizer.Config):         # diff_sugar_rasterizer_temporal.py:58-62 / _normal.py:56-60
        debug: bool = False
        invert_bg_prob: float = 1.0
        back_ground_color: Tuple[float, float, float] = (1, 1, 1)

    cfg: Config

    def configure(self, geometry, material=None, background=None) -> None:
        super().configure(geometry, material, background)
        self._impl = DiffGaussianBatchRenderer(geometry, tuple(self.cfg.back_ground_color), training=True)
        self.background_tensor = torch.tensor(tuple(self.cfg.back_ground_color), dtype=torch.float32, device=self.device)

    @property
    def capacity(self):
        return self._impl.capacity

    @capacity.setter
    def capacity(self, value):
        """None (default): exact sizing with one read-back per batch; an integer keeps the step free of host
        synchronisation (CUDA-graph capture) — see ``trainstep.GraphedDynamicStageStep``."""
        self._impl.capacity = value

    @property
    def last_state(self):
        return self._impl.last_state

    def batch_forward(self, batch: Dict[str, Any], **kwargs) -> Dict[str, Any]:
        self._impl.training = self.training          # eval inverts the background (temporal.py:96-103)
        return self._impl.batch_forward(batch, **kwargs)

    def forward(self, viewpoint_camera=None, bg_color=None, **batch) -> Dict[str, Any]:
        """Single-view form of DiffGaussian.forward (temporal.py:81-239): the batch entry ``batch_idx`` of ``batch``."""
        i = int(batch.get("batch_idx", 0))
        one = {k: (v[i:i + 1] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] > i else v) for k, v in batch.items()}
        out = self.batch_forward(one)
        return {"render": out["comp_rgb"][0].permute(2, 0, 1), "normal": out["comp_normal"][0].permute(2, 0, 1),
                "depth": out["comp_depth"][0].permute(2, 0, 1), "mask": out["comp_mask"][0].permute(2, 0, 1),
                "viewspace_points": out["viewspace_points"], "visibility_filter": out["visibility_filter"][0],
                "radii": out["radii"][0],
                **({"normal_from_dist": out["comp_normal_from_dist"][0].permute(2, 0, 1)} if "comp_normal_from_dist" in out else {})}


@_provides("diff-sugar-rasterizer-temporal")
class DiffGaussian(_BatchedRasterizer):
    pass


@_provides("diff-sugar-rasterizer-normal")
class DiffSuGaR(_BatchedRasterizer):
    pass


# ------------------------------------------------------------------------------------------------------------------
# registration
# ------------------------------------------------------------------------------------------------------------------
def install(override: bool = True) -> Dict[str, type]:
    """Puts the classes above into threestudio's registry under the reference's names.  Works whichever of the two
    plugins is imported first: names the reference already registered are replaced, and its later ``@register`` of one of
    these names is ignored instead of raising "Names of extensions conflict" (threestudio/__init__.py:5-15)."""
    if not HAVE_THREESTUDIO:
        return dict(REGISTRY)
    import threestudio as ts
    for name, cls in REGISTRY.items():
        if override or name not in ts.__modules__:
            ts.__modules__[name] = cls
    if not getattr(ts.register, "_dm4d_wrapped", False):
        original = ts.register

        def register(name):
            if name in REGISTRY and name in ts.__modules__:
                return lambda cls: cls          # keep the B200 class; the reference's class object stays importable
            return original(name)
        register._dm4d_wrapped = True
        ts.register = register
    return dict(REGISTRY)
