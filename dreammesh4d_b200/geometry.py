"""Host-side mirror of the reference's geometry plugin interface for the hot path.

``DynamicSuGaRGeometry`` exposes the getters the renderer and the system call on
``DynamicSuGaRModel`` / ``SuGaRModel`` (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py,
geometry/sugar.py) with the same names, argument meaning and tensor layouts, but evaluates ALL
timestamps of a step with the fused kernels of libdm4d.so instead of ~150 eager ops per view.
State tensors keep the reference's attribute names so its checkpoints map 1:1.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

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


class DynamicSuGaRGeometry(nn.Module):
    """Surface-bound Gaussians on a mesh deformed by a sparse control graph.

    ``deformation(node_xyz [M,3], timestamps [T]) -> (trans [T,M,3], rot_delta [T,M,4], strain [T,M,6],
    opacity_delta [T,M,1])`` is the (PyTorch) deformation network — DeformationNetwork.forward_dynamic_delta
    in the reference (geometry/deformation.py:538-539), evaluated at ``2 t - 1`` by the caller there
    (dynamic_sugar.py:431); pass any callable, e.g. ``dreammesh4d_b200.deformation.HexPlaneDeformation``.
    """

    def __init__(self, scene: SugarScene, graph: DeformGraph, deformation: Optional[Callable] = None,
                 skinning_method: str = "hybrid", static_learnable: bool = False):
        super().__init__()
        self.g = scene.g
        self.skinning_method = skinning_method
        self.thickness = float(scene.thickness)
        rg = static_learnable
        # reference attribute names
        self._points = nn.Parameter(scene.verts.clone().float(), requires_grad=rg)
        self._scales = nn.Parameter(scene.log_scales.clone().float(), requires_grad=rg)
        self._quaternions = nn.Parameter(scene.complex_rot.clone().float(), requires_grad=rg)
        self.all_densities = nn.Parameter(scene.densities.clone().float(), requires_grad=rg)
        self._sh_coordinates_dc = nn.Parameter(scene.sh_dc.clone().float(), requires_grad=rg)
        self.register_buffer("_surface_mesh_faces", scene.faces.clone().long())
        self.register_buffer("_faces_i32", scene.faces.clone().int(), persistent=False)
        self.register_buffer("surface_triangle_bary_coords", scene.bary.clone().float()[..., None])
        # deformation graph: buffers here (the reference keeps them as plain attributes)
        self.register_buffer("_deform_graph_node_xyz", graph.node_xyz.clone().float())
        self.register_buffer("_xyz_neighbor_node_idx", graph.nbr_idx.clone().long())
        self.register_buffer("_nbr_idx_i32", graph.nbr_idx.clone().int(), persistent=False)
        self.register_buffer("_xyz_neighbor_nodes_weights", graph.nbr_w.clone().float())
        self._deformation = deformation
        self._step_cache: Dict[str, torch.Tensor] = {}
        self._timed: Optional[dict] = None

    # ---- static getters (sugar.py:440-548,640-648) ------------------------------------------------
    @property
    def n_gaussians(self) -> int:
        return self._surface_mesh_faces.shape[0] * self.g

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

    # ---- dynamic path ------------------------------------------------------------------------------
    def get_timed_dg_attributes(self, timestamp: torch.Tensor):
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
        self._timed = {"timestamp": timestamp, "means3D": means, "rotations": rots, "normals": normals,
                       "verts": verts, "vert_rot": vrot}
        return self._timed

    def _timed_index(self, timestamp):
        if self._timed is None:
            raise RuntimeError("call deform(timestamps) first (the batched renderer does)")
        if timestamp.ndim == 0:
            timestamp = timestamp[None]
        # device-side lookup of the cached row(s); no host sync
        return (self._timed["timestamp"][None, :] == timestamp[:, None]).float().argmax(dim=1)

    def get_timed_gs_all_single_time(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:708-724 — (means3D, scales, rotations, opacity, colors_precomp) of one view."""
        if self._timed is None or self._timed["timestamp"].shape[0] != 1:
            self.deform(timestamp.reshape(1))
        return (self._timed["means3D"][0], self.get_scaling, self._timed["rotations"][0], self.get_opacity,
                self.get_points_rgb())

    def get_timed_gs_normals(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:357-364 — [N_t, P, 3]."""
        return self._timed["normals"][self._timed_index(timestamp)]

    def get_timed_vertex_xyz(self, timestamp=None, frame_idx=None):
        """dynamic_sugar.py:281-301 — [N_t, V, 3]."""
        return self._timed["verts"][self._timed_index(timestamp)]

    def get_timed_vertex_rotation(self, timestamp=None, frame_idx=None, return_matrix: bool = False):
        """dynamic_sugar.py:303-327 — xyzw quaternions [N_t,V,4] or matrices [N_t,V,3,3]."""
        q = self._timed["vert_rot"][self._timed_index(timestamp)]
        if not return_matrix:
            return q
        x, y, z, w = q.unbind(-1)
        return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                            2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                            2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1).reshape(*q.shape[:-1], 3, 3)

    def update_step(self, epoch: int = 0, global_step: int = 0, on_load_weights: bool = False):
        """dynamic_sugar.py:863-873 — clears the per-step caches."""
        self._step_cache = {}
        self._timed = None
