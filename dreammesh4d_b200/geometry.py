"""Host-side mirror of the reference's geometry plugin interface for the hot path.

``DynamicSuGaRGeometry`` exposes the getters the renderer and the system call on
``DynamicSuGaRModel`` / ``SuGaRModel`` (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py,
geometry/sugar.py) with the same names, argument meaning and tensor layouts, but evaluates ALL
timestamps of a step with the fused kernels of libdm4d.so instead of ~150 eager ops per view.
State tensors keep the reference's attribute names so its checkpoints map 1:1.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import skinning
from .synthetic import C0, DeformGraph, SugarScene


def strain_tensor_to_matrix(strain: torch.Tensor) -> torch.Tensor:
    """dynamic_sugar.py:29-39 — I + symmetric strain, [..., 6] -> [..., 3, 3]."""
    d = strain[..., :3]
    o = strain[..., 3:]
    one = torch.ones_like(d[..., 0])
    rows = [one + d[..., 0], o[..., 0], o[..., 1],
            o[..., 0], one + d[..., 1], o[..., 2],
            o[..., 1], o[..., 2], one + d[..., 2]]
    return torch.stack(rows, dim=-1).reshape(*strain.shape[:-1], 3, 3)


def activate_node_deltas(trans, rot_delta, strain6, opacity_delta):
    """dynamic_sugar.py:445-465: reshape/activate the deformation network's raw outputs.
    rot = normalize(delta + (0,0,0,1)) in xyzw, scale = I + sym(strain), opacity = sigmoid."""
    ident = torch.zeros_like(rot_delta)
    ident[..., 3] = 1.0
    rot = F.normalize(rot_delta + ident, dim=-1)
    scale = None if strain6 is None else strain_tensor_to_matrix(strain6)
    opacity = None if opacity_delta is None else torch.sigmoid(opacity_delta)
    return trans, rot, scale, opacity


def deformed_gaussian_scales(node_scale: torch.Tensor, node_opacity: Optional[torch.Tensor], nbr_idx: torch.Tensor,
                             nbr_w: torch.Tensor, faces: torch.Tensor, bary: torch.Tensor, scaling: torch.Tensor,
                             method: str) -> torch.Tensor:
    """``d_scale=True`` (the reference's class default, off in the shipped YAML): per-Gaussian scale vectors [T,P,3] =
    (barycentric blend over the face's vertices of the blended node scale matrices) x static scaling vector.
    Vertex matrices (dynamic_sugar.py:595-612): LBS ``sum_k w_k S_k``; hybrid ``sum_k w_k o_k S_k + (1 - lambda) I`` with
    ``lambda = min(1, sum_k w_k o_k + 0.4)`` (:571-577); Gaussian scales (:698-704).  Plain tensor ops (autograd gives
    the gradients to the node attributes): an optional branch, not part of the fused skinning kernels."""
    if method == "dqs":
        raise ValueError("d_scale is undefined for skinning_method='dqs' (the reference produces no vertex scale there)")
    Sn = node_scale[:, nbr_idx]                                            # [T,V,K,3,3]
    w = nbr_w[None, ..., None, None]
    if method == "lbs":
        Sv = (w * Sn).sum(dim=-3)
    else:
        on = node_opacity.reshape(node_opacity.shape[0], -1, 1)[:, nbr_idx]   # [T,V,K,1]
        lam = ((nbr_w[None, ..., None] * on).sum(dim=-2) + 0.4).clamp(max=1.0)
        Sv = (w * on[..., None] * Sn).sum(dim=-3) + (1.0 - lam)[..., None] * torch.eye(3, dtype=Sn.dtype, device=Sn.device)
    T, F_, g = Sv.shape[0], faces.shape[0], bary.shape[0]
    Sg = (bary[None, None, :, :, None, None] * Sv[:, faces][:, :, None]).sum(dim=-3)       # [T,F,g,3,3]
    return (Sg.reshape(T, F_ * g, 3, 3) @ scaling[None, ..., None]).squeeze(-1)


def exp_interp(value, step: int) -> float:
    """threestudio ``C(value, 0, step, interpolation="exp")`` (threestudio/utils/misc.py:66-101) for the learning-rate
    schedules of sugar.py:387-404 / dynamic_sugar.py:230-279: a number, or [start_step, v0, v1, end_step]."""
    if isinstance(value, (int, float)):
        return float(value)
    v = list(value)
    if len(v) == 3:
        v = [0] + v
    start, v0, v1, end = v
    t = max(min(1.0, (step - start) / (end - start)), 0.0)
    return math.exp(math.log(v0) * (1 - t) + math.log(v1) * t)


class TimedMeshes:
    """What ``get_timed_surface_mesh`` hands out when pytorch3d is not installed: the fields the hot path's consumers
    read (``verts_padded`` / ``faces_padded``), enough for ``arap.mesh_normal_consistency`` and the exporters."""

    def __init__(self, verts: torch.Tensor, faces: torch.Tensor, vertex_colors: Optional[torch.Tensor]):
        self._v, self._f, self._c = verts, faces, vertex_colors

    def verts_padded(self):
        return self._v

    def verts_list(self):
        return list(self._v.unbind(0))

    def faces_list(self):
        return [self._f] * self._v.shape[0]

    def faces_padded(self):
        return self._f[None].expand(self._v.shape[0], -1, -1)

    def __len__(self):
        return self._v.shape[0]


class SuGaRState:
    """State + getters of the surface-bound Gaussian model, shared by the stand-alone ``DynamicSuGaRGeometry`` and the
    registered threestudio classes of ``plugin.py`` (a mix-in: the host class is an ``nn.Module``).  Attribute names
    are the reference's so its checkpoints map 1:1."""

    # ---- construction ------------------------------------------------------------------------------------------
    def _install_state(self, scene: SugarScene, graph: Optional[DeformGraph], deformation: Optional[Callable],
                       skinning_method: str = "hybrid", static_learnable: bool = False, sh_levels: int = 1,
                       learn: Optional[Dict[str, bool]] = None) -> None:
        self.g = scene.g
        self.skinning_method = skinning_method
        self.d_scale = False                      # see deformed_gaussian_scales
        self.thickness = float(scene.thickness)
        lr = dict(points=static_learnable, scales=static_learnable, quaternions=static_learnable,
                  densities=static_learnable, sh=static_learnable)
        lr.update(learn or {})
        # reference attribute names
        self.surface_mesh_thickness = nn.Parameter(torch.tensor(self.thickness), requires_grad=False)   # sugar.py:192-196
        self._points = nn.Parameter(scene.verts.clone().float(), requires_grad=lr["points"])
        self._scales = nn.Parameter(scene.log_scales.clone().float(), requires_grad=lr["scales"])
        self._quaternions = nn.Parameter(scene.complex_rot.clone().float(), requires_grad=lr["quaternions"])
        self.all_densities = nn.Parameter(scene.densities.clone().float(), requires_grad=lr["densities"])
        self._sh_coordinates_dc = nn.Parameter(scene.sh_dc.clone().float(), requires_grad=lr["sh"])
        self._sh_coordinates_rest = nn.Parameter(torch.zeros(scene.sh_dc.shape[0], sh_levels ** 2 - 1, 3), requires_grad=lr["sh"])
        self.register_buffer("_surface_mesh_faces", scene.faces.clone().long())
        self.register_buffer("_faces_i32", scene.faces.clone().int(), persistent=False)
        self.register_buffer("surface_triangle_bary_coords", scene.bary.clone().float()[..., None], persistent=False)
        vc = getattr(scene, "vertex_colors", None)
        self.register_buffer("_vertex_colors", None if vc is None else vc.clone().float(), persistent=False)
        self._deformation = deformation
        self._step_cache: Dict[str, torch.Tensor] = {}
        self._timed: Optional[dict] = None
        if graph is not None:
            self._install_graph(graph)

    def _install_graph(self, graph: DeformGraph) -> None:
        """Deformation graph as BUFFERS (persistent: a checkpoint is self-contained; the reference keeps them as plain
        attributes and re-samples the nodes on reload)."""
        for name, t, persistent in (("_deform_graph_node_xyz", graph.node_xyz.clone().float(), True),
                                    ("_xyz_neighbor_node_idx", graph.nbr_idx.clone().long(), True),
                                    ("_nbr_idx_i32", graph.nbr_idx.clone().int(), False),
                                    ("_xyz_neighbor_nodes_weights", graph.nbr_w.clone().float(), True)):
            if name in self._buffers:
                self._buffers[name] = t.to(self._points.device)
            else:
                self.register_buffer(name, t, persistent=persistent)

    # ---- static getters (sugar.py:418-548,640-648) -----------------------------------------------------------------
    @property
    def n_gaussians(self) -> int:
        return self._surface_mesh_faces.shape[0] * self.g

    n_points = n_gaussians

    @property
    def n_verts(self) -> int:
        return self._points.shape[0]

    @property
    def n_faces(self) -> int:
        return self._surface_mesh_faces.shape[0]

    @property
    def get_xyz_verts(self):
        return self._points

    @property
    def get_faces(self):
        return self._surface_mesh_faces

    @property
    def get_xyz(self):
        fv = self._points[self._surface_mesh_faces]
        return (fv[:, None] * self.surface_triangle_bary_coords[None]).sum(dim=-2).reshape(-1, 3)

    @property
    def get_scaling(self):
        return torch.cat([torch.full_like(self._scales[:, :1], self.thickness), torch.exp(self._scales)], dim=-1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self.all_densities.view(-1, 1))

    def get_points_rgb(self):
        return (self._sh_coordinates_dc * C0 + 0.5).view(-1, 3)

    def _rest_frames(self):
        if "rest_q" not in self._step_cache:
            q, n = skinning.sugar_rest_frames(self._points, self._faces_i32, self._quaternions, self.g)
            self._step_cache["rest_q"], self._step_cache["rest_n"] = q, n
        return self._step_cache["rest_q"], self._step_cache["rest_n"]

    @property
    def get_rotation(self):
        return self._rest_frames()[0]

    @property
    def get_gs_normals(self):
        return self._rest_frames()[1]

    # ---- dynamic path ------------------------------------------------------------------------------------------------
    def get_timed_dg_attributes(self, timestamp: torch.Tensor, frame_idx=None):
        """Activated control-node attributes for timestamps [T] (dynamic_sugar.py:367-465)."""
        if self._deformation is None:
            raise RuntimeError("no deformation network attached")
        return activate_node_deltas(*self._deformation(self._deform_graph_node_xyz, timestamp))

    def deform(self, timestamp: torch.Tensor, node_attrs=None) -> dict:
        """Deforms mesh + Gaussians for all timestamps [T] of a step in one launch sequence and caches
        the result until ``update_step``.  Returns dict(means3D [T,P,3], rotations [T,P,4] wxyz,
        normals [T,P,3], verts [T,V,3], vert_rot [T,V,4] xyzw)."""
        trans, rot, scale, opac = node_attrs if node_attrs is not None else self.get_timed_dg_attributes(timestamp)
        means, rots, normals, verts, vrot = skinning.skin_gaussians(
            trans, rot, scale, opac, self._points, self._faces_i32, self._nbr_idx_i32,
            self._xyz_neighbor_nodes_weights, self.surface_triangle_bary_coords[..., 0], self.get_rotation,
            method=self.skinning_method)
        self._timed = {"timestamp": timestamp, "key": (timestamp.data_ptr(), timestamp._version, tuple(timestamp.shape)),
                       "means3D": means, "rotations": rots, "normals": normals, "verts": verts, "vert_rot": vrot}
        if self.d_scale:
            self._timed["scales"] = deformed_gaussian_scales(
                scale, opac, self._xyz_neighbor_node_idx, self._xyz_neighbor_nodes_weights, self._surface_mesh_faces,
                self.surface_triangle_bary_coords[..., 0], self.get_scaling, self.skinning_method)
        return self._timed

    def _timed_for(self, timestamp: torch.Tensor) -> dict:
        """The cached deformation if it was computed for this very timestamp tensor, else a fresh one.  Identity, not
        values, decides: comparing values would need a host synchronisation."""
        if timestamp.ndim == 0:
            timestamp = timestamp[None]
        if self._timed is not None and self._timed["key"] == (timestamp.data_ptr(), timestamp._version, tuple(timestamp.shape)):
            return self._timed
        return self.deform(timestamp)

    def _timed_index(self, timestamp):
        if self._timed is None:
            raise RuntimeError("call deform(timestamps) first (the batched renderer does)")
        if timestamp.ndim == 0:
            timestamp = timestamp[None]
        # device-side lookup of the cached row(s); no host sync
        return (self._timed["timestamp"][None, :] == timestamp[:, None]).float().argmax(dim=1)

    def get_timed_gs_all_single_time(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:708-724 — (means3D, scales, rotations, opacity, colors_precomp) of one view."""
        t = self._timed_for(timestamp.reshape(1))
        scales = t["scales"][0] if "scales" in t else self.get_scaling
        return t["means3D"][0], scales, t["rotations"][0], self.get_opacity, self.get_points_rgb()

    def get_timed_gs_normals(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:357-364 — [N_t, P, 3]."""
        return self._timed_for(timestamp)["normals"]

    def get_timed_vertex_xyz(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:281-301 — [N_t, V, 3]."""
        return self._timed_for(timestamp)["verts"]

    def get_timed_vertex_rotation(self, timestamp=None, frame_idx=None, return_matrix: bool = False):
        """dynamic_sugar.py:303-327 — xyzw quaternions [N_t,V,4] or matrices [N_t,V,3,3]."""
        q = self._timed_for(timestamp)["vert_rot"]
        if not return_matrix:
            return q
        x, y, z, w = q.unbind(-1)
        return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(*q.shape[:-1], 3, 3)

    def get_timed_surface_mesh(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:330-345 — the deformed meshes of the given timestamps: a pytorch3d ``Meshes`` with vertex
        colour textures when pytorch3d is installed (what the system hands to ``mesh_normal_consistency``,
        sugar_4dgen.py:218-225), else ``TimedMeshes``."""
        verts = self.get_timed_vertex_xyz(timestamp, frame_idx)
        try:
            from pytorch3d.renderer import TexturesVertex
            from pytorch3d.structures import Meshes
        except ImportError:
            return TimedMeshes(verts, self._surface_mesh_faces, self._vertex_colors)
        n_t = verts.shape[0]
        tex = None if self._vertex_colors is None else \
            TexturesVertex(verts_features=self._vertex_colors.clamp(0, 1)[None].expand(n_t, -1, -1))
        return Meshes(verts=verts, faces=self._surface_mesh_faces[None].expand(n_t, -1, -1), textures=tex)

    @property
    def surface_mesh(self):
        """sugar.py:578-586 — the rest-pose mesh (static stage: mesh regularisers sugar_static.py:244-253; export)."""
        try:
            from pytorch3d.renderer import TexturesVertex
            from pytorch3d.structures import Meshes
        except ImportError:
            return TimedMeshes(self._points[None], self._surface_mesh_faces, self._vertex_colors)
        tex = None if self._vertex_colors is None else TexturesVertex(verts_features=self._vertex_colors.clamp(0, 1)[None])
        return Meshes(verts=[self._points], faces=[self._surface_mesh_faces], textures=tex)

    @property
    def _deformed_vert_positions(self) -> Dict[int, torch.Tensor]:
        """dynamic_sugar.py:289-297 keeps a dict keyed by (timestamp, frame); the system only iterates its values
        (sugar_4dgen.py:294-297)."""
        return {} if self._timed is None else {i: v for i, v in enumerate(self._timed["verts"].unbind(0))}

    def update_step(self, epoch: int = 0, global_step: int = 0, on_load_weights: bool = False):
        """dynamic_sugar.py:863-873 — clears the per-step caches."""
        self._step_cache = {}
        self._timed = None

    # ---- optimizer groups (sugar.py:327-416, dynamic_sugar.py:168-279) ----------------------------------------------
    def _static_groups(self, cfg) -> list:
        s = float(cfg.spatial_lr_scale)
        groups = []
        if self._points.requires_grad:
            groups.append({"params": [self._points], "lr": exp_interp(cfg.position_lr, 0) * s, "name": "points"})
        if self._sh_coordinates_dc.requires_grad:
            groups.append({"params": [self._sh_coordinates_dc], "lr": exp_interp(cfg.feature_lr, 0), "name": "f_dc"})
            groups.append({"params": [self._sh_coordinates_rest], "lr": exp_interp(cfg.feature_lr, 0) / 20.0, "name": "f_rest"})
        if self.all_densities.requires_grad:
            groups.append({"params": [self.all_densities], "lr": exp_interp(cfg.opacity_lr, 0), "name": "all_densities"})
        if self._scales.requires_grad:
            groups.append({"params": [self._scales], "lr": exp_interp(cfg.scaling_lr, 0), "name": "scales"})
            groups.append({"params": [self._quaternions], "lr": exp_interp(cfg.rotation_lr, 0), "name": "quaternions"})
        return groups

    def _dynamic_groups(self, cfg) -> list:
        net = self._deformation
        s = float(cfg.spatial_lr_scale)
        grid = [p for n, p in net.named_parameters() if "grid" in n and p.requires_grad]       # deformation.py get_grid_parameters
        mlp = [p for n, p in net.named_parameters() if "grid" not in n and p.requires_grad]    # get_mlp_parameters
        return [{"params": mlp, "lr": exp_interp(cfg.deformation_lr, 0) * s, "name": "deformation"},
                {"params": grid, "lr": exp_interp(cfg.grid_lr, 0) * s, "name": "grid"}]

    def _set_optimizer(self, groups: list) -> None:
        self.optimize_list = groups
        self.optimize_params = [g["name"] for g in groups]
        self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15) if groups else None

    def update_learning_rate(self, iteration: int) -> None:
        """sugar.py:387-404 + dynamic_sugar.py:230-279: exponential interpolation of the scheduled rates."""
        cfg = getattr(self, "cfg", None)
        if cfg is None or getattr(self, "optimizer", None) is None:
            return
        s = float(cfg.spatial_lr_scale)
        rules = {"points": lambda: exp_interp(cfg.position_lr, iteration) * s,
                 "f_dc": lambda: exp_interp(cfg.feature_lr, iteration),
                 "f_rest": lambda: exp_interp(cfg.feature_lr, iteration) / 20.0,
                 "grid": lambda: exp_interp(cfg.grid_lr, iteration) * s,
                 "deformation": lambda: exp_interp(cfg.deformation_lr, iteration) * s}
        for group in self.optimizer.param_groups:
            rule = rules.get(group.get("name"))
            if rule is not None:
                group["lr"] = rule()

    def merge_optimizer(self, net_optimizer):
        """sugar.py:406-416 — geometry groups + the system's groups in one AdamW(betas=(0.9, 0.99), eps=1e-15)."""
        groups = list(self.optimize_list)
        for param in net_optimizer.param_groups:
            groups.append({"params": param["params"], "lr": param["lr"]})
        self.optimize_list = groups
        self.optimizer = torch.optim.AdamW(groups, lr=0.0, betas=(0.9, 0.99), eps=1e-15)
        return self.optimizer


class DynamicSuGaRGeometry(SuGaRState, nn.Module):
    """Surface-bound Gaussians on a mesh deformed by a sparse control graph (stand-alone form; the registered
    threestudio classes live in ``plugin.py``).

    ``deformation(node_xyz [M,3], timestamps [T]) -> (trans [T,M,3], rot_delta [T,M,4], strain [T,M,6],
    opacity_delta [T,M,1])`` is the (PyTorch) deformation network — DeformationNetwork.forward_dynamic_delta
    in the reference (geometry/deformation.py:538-539), evaluated at ``2 t - 1`` by the caller there
    (dynamic_sugar.py:431); pass any callable, e.g. ``dreammesh4d_b200.deformation.HexPlaneDeformation``.
    """

    def __init__(self, scene: SugarScene, graph: DeformGraph, deformation: Optional[Callable] = None,
                 skinning_method: str = "hybrid", static_learnable: bool = False):
        super().__init__()
        self._install_state(scene, graph, deformation, skinning_method, static_learnable)
