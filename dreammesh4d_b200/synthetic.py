"""Synthetic meshes / surface-bound Gaussians / deformation graphs / cameras of the sizes named in
BASELINE.json (distributions and seeds: SURVEY.md §8d).  Pure torch, device-agnostic; used by
bench.py, __graft_entry__.smoke() and the tests.  No reference data is needed or read.

Reference conventions reproduced here:
  * barycentric tables and circle radii: custom/threestudio-dreammesh4d/geometry/sugar.py:235-276
  * initial scales / complex rotations / thickness: sugar.py:191,301-325
  * deformation-graph weights (1 - d_k/d_{K+1})^2, row-normalised: geometry/dynamic_sugar.py:845,859-861
  * orbit cameras: configs/sugar_dynamic_dg.yaml:15-31 (distance 3.8, fovy 20 deg, look-at origin, up +z)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

C0 = 0.28209479177387814   # geometry/gaussian_base.py:32

BARY = {
    1: [[1 / 3, 1 / 3, 1 / 3]],
    3: [[1 / 2, 1 / 4, 1 / 4], [1 / 4, 1 / 2, 1 / 4], [1 / 4, 1 / 4, 1 / 2]],
    4: [[1 / 3, 1 / 3, 1 / 3], [2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3]],
    6: [[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3],
        [1 / 6, 5 / 12, 5 / 12], [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]],
}


def circle_radius(g: int, s: float = 1.3) -> float:
    """surface_triangle_circle_radius (sugar.py:237,245,255,266)."""
    return {1: 1.0 / 2.0 / math.sqrt(3.0), 3: 1.0 / 2.0 / (math.sqrt(3.0) + 1.0), 4: 1.0 / (4.0 * math.sqrt(3.0)),
            6: 1.0 / (4.0 + 2.0 * math.sqrt(3.0))}[g] * s


def uv_sphere(n_faces: int, radius: float = 0.5):
    """Closed UV sphere with exactly ``n_faces`` triangles (n_faces = 2 * n_lon * (n_lat - 1))."""
    if n_faces % 2:
        raise ValueError("n_faces must be even")
    half = n_faces // 2
    best = None
    for n_lon in range(3, half + 1):
        if half % n_lon:
            continue
        rings = half // n_lon            # n_lat - 1
        if rings < 2:
            break
        score = abs(n_lon / (2.0 * rings) - 1.0)
        if best is None or score < best[0]:
            best = (score, n_lon, rings)
    if best is None:
        raise ValueError(f"cannot build a UV sphere with {n_faces} faces")
    _, n_lon, rings = best
    n_lat = rings + 1                    # latitude bands incl. the two polar fans
    theta = torch.linspace(0, math.pi, n_lat + 1)[1:-1]                 # interior rings
    phi = torch.arange(n_lon) * (2 * math.pi / n_lon)
    ring = torch.stack([torch.sin(theta)[:, None] * torch.cos(phi)[None], torch.sin(theta)[:, None] * torch.sin(phi)[None],
                        torch.cos(theta)[:, None].expand(-1, n_lon)], dim=-1).reshape(-1, 3)
    verts = torch.cat([torch.tensor([[0.0, 0.0, 1.0]]), ring, torch.tensor([[0.0, 0.0, -1.0]])]) * radius
    n_rings = n_lat - 1
    idx = lambda r, c: 1 + r * n_lon + (c % n_lon)
    faces = []
    c = torch.arange(n_lon)
    faces.append(torch.stack([torch.zeros(n_lon, dtype=torch.long), idx(0, c), idx(0, c + 1)], dim=1))
    for r in range(n_rings - 1):
        a, b, cc, d = idx(r, c), idx(r + 1, c), idx(r + 1, c + 1), idx(r, c + 1)
        faces.append(torch.stack([a, b, cc], dim=1))
        faces.append(torch.stack([a, cc, d], dim=1))
    south = 1 + n_rings * n_lon
    faces.append(torch.stack([idx(n_rings - 1, c), torch.full((n_lon,), south, dtype=torch.long), idx(n_rings - 1, c + 1)], dim=1))
    faces = torch.cat(faces)
    assert faces.shape[0] == n_faces, (faces.shape, n_faces)
    return verts.float(), faces.long()


@dataclass
class SugarScene:
    """Static SuGaR state of a mesh-bound Gaussian cloud (checkpoint schema: SURVEY.md Appendix D)."""
    verts: torch.Tensor          # [V,3]   _points
    faces: torch.Tensor          # [F,3]   _surface_mesh_faces (int64)
    bary: torch.Tensor           # [g,3]   surface_triangle_bary_coords
    log_scales: torch.Tensor     # [P,2]   _scales
    complex_rot: torch.Tensor    # [P,2]   _quaternions
    densities: torch.Tensor      # [P,1]   all_densities (pre-sigmoid)
    sh_dc: torch.Tensor          # [P,1,3] _sh_coordinates_dc
    thickness: float             # surface_mesh_thickness = spatial_extent / 1e6
    g: int

    @property
    def n_gaussians(self) -> int:
        return self.faces.shape[0] * self.g

    def to(self, device):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


def make_sugar_scene(n_faces: int, g: int = 3, radius: float = 0.5, init_scale_s: float = 1.3, opacity: float = 0.9,
                     spatial_extent: float = 1.0) -> SugarScene:
    verts, faces = uv_sphere(n_faces, radius)
    bary = torch.tensor(BARY[g], dtype=torch.float32)
    fv = verts[faces]                                            # [F,3,3]
    edge = (fv - fv[:, [1, 2, 0]]).norm(dim=-1).min(dim=-1)[0]    # sugar.py:309
    scales = (edge * circle_radius(g, init_scale_s)).clamp_min(1e-7)
    log_scales = scales.log()[:, None, None].expand(-1, g, 2).reshape(-1, 2).contiguous()
    P = n_faces * g
    complex_rot = torch.zeros(P, 2)
    complex_rot[:, 0] = 1.0
    densities = torch.full((P, 1), math.log(opacity / (1 - opacity)))
    normals = torch.nn.functional.normalize(verts, dim=-1)
    vcol = 0.5 + 0.5 * normals
    col = (vcol[faces][:, None] * bary[None, :, :, None]).sum(dim=-2).reshape(-1, 3)   # sugar.py:214-218
    sh_dc = ((col - 0.5) / C0)[:, None, :].contiguous()
    return SugarScene(verts, faces, bary, log_scales, complex_rot, densities, sh_dc, spatial_extent / 1_000_000, g)


@dataclass
class DeformGraph:
    node_xyz: torch.Tensor       # [M,3]  _deform_graph_node_xyz
    nbr_idx: torch.Tensor        # [V,K]  _xyz_neighbor_node_idx (int64)
    nbr_w: torch.Tensor          # [V,K]  _xyz_neighbor_nodes_weights

    def to(self, device):
        self.node_xyz, self.nbr_idx, self.nbr_w = (t.to(device) for t in (self.node_xyz, self.nbr_idx, self.nbr_w))
        return self


def farthest_point_sample(x: torch.Tensor, m: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    n = x.shape[0]
    sel = torch.empty(m, dtype=torch.long)
    sel[0] = torch.randint(n, (1,), generator=g).item()
    d = torch.full((n,), float("inf"))
    for i in range(1, m):
        d = torch.minimum(d, (x - x[sel[i - 1]]).pow(2).sum(-1))
        sel[i] = torch.argmax(d)
    return sel


def make_deform_graph(verts: torch.Tensor, M: int, K: int, seed: int = 0) -> DeformGraph:
    verts = verts.cpu()
    nodes = verts[farthest_point_sample(verts, M, seed)]
    idxs, ws = [], []
    for chunk in verts.split(16384):
        dist = torch.cdist(chunk, nodes)
        dk, ik = dist.topk(K + 1, dim=1, largest=False)
        w = (1 - dk[:, :K] / dk[:, K:K + 1].clamp_min(1e-12)) ** 2           # dynamic_sugar.py:845
        w = w / w.sum(dim=-1, keepdim=True).clamp_min(1e-12)                 # :859-861
        idxs.append(ik[:, :K])
        ws.append(w)
    return DeformGraph(nodes.float(), torch.cat(idxs).long(), torch.cat(ws).float())


def random_node_attrs(n_t: int, M: int, seed: int = 1):
    """Activated node attributes as produced by _get_timed_dg_attributes (dynamic_sugar.py:408-465):
    trans [n_t,M,3], rot xyzw unit [n_t,M,4], scale = I + sym(strain) [n_t,M,3,3], opacity in (0,1) [n_t,M,1]."""
    g = torch.Generator().manual_seed(seed)
    trans = torch.randn(n_t, M, 3, generator=g) * 0.02
    rot = torch.cat([torch.randn(n_t, M, 3, generator=g) * 0.1, torch.ones(n_t, M, 1)], dim=-1)
    rot = torch.nn.functional.normalize(rot, dim=-1)
    strain = torch.randn(n_t, M, 6, generator=g) * 0.02
    scale = torch.eye(3).expand(n_t, M, 3, 3).clone()
    scale[..., 0, 0] += strain[..., 0]; scale[..., 1, 1] += strain[..., 1]; scale[..., 2, 2] += strain[..., 2]
    scale[..., 0, 1] += strain[..., 3]; scale[..., 0, 2] += strain[..., 4]; scale[..., 1, 2] += strain[..., 5]
    scale[..., 1, 0] += strain[..., 3]; scale[..., 2, 0] += strain[..., 4]; scale[..., 2, 1] += strain[..., 5]
    opacity = torch.sigmoid(torch.randn(n_t, M, 1, generator=g))
    return trans, rot, scale, opacity


def orbit_c2w(elevation_deg: torch.Tensor, azimuth_deg: torch.Tensor, distance: float = 3.8) -> torch.Tensor:
    """Camera-to-world [B,4,4] in the threestudio convention (camera looks down -z, +y up; world up +z),
    as built by data/uncond.py / data/temporal_image.py."""
    e, a = torch.deg2rad(elevation_deg.double()), torch.deg2rad(azimuth_deg.double())
    pos = torch.stack([distance * torch.cos(e) * torch.cos(a), distance * torch.cos(e) * torch.sin(a),
                       distance * torch.sin(e)], dim=-1)
    lookat = torch.nn.functional.normalize(-pos, dim=-1)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64).expand_as(pos)
    right = torch.nn.functional.normalize(torch.cross(lookat, up, dim=-1), dim=-1)
    up2 = torch.nn.functional.normalize(torch.cross(right, lookat, dim=-1), dim=-1)
    c2w = torch.eye(4, dtype=torch.float64).repeat(pos.shape[0], 1, 1)
    c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = right, up2, -lookat, pos
    return c2w.float()


def random_orbit_cameras(B: int, seed: int = 2, distance: float = 3.8, fovy_deg: float = 20.0):
    """elevation U(-10,80), azimuth U(-180,180). Returns (c2w [B,4,4], fovy [B] rad)."""
    g = torch.Generator().manual_seed(seed)
    elev = torch.rand(B, generator=g) * 90.0 - 10.0
    azim = torch.rand(B, generator=g) * 360.0 - 180.0
    return orbit_c2w(elev, azim, distance), torch.full((B,), math.radians(fovy_deg))


def random_gaussians(P: int, seed: int = 0):
    """Config C4: free Gaussians in a ball of radius 0.5."""
    g = torch.Generator().manual_seed(seed)
    d = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
    means = d * (torch.rand(P, 1, generator=g) ** (1.0 / 3.0)) * 0.5
    scales = torch.exp(torch.rand(P, 3, generator=g) * (math.log(1e-2) - math.log(1e-3)) + math.log(1e-3))
    rots = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1)
    opac = torch.rand(P, 1, generator=g) * 0.94 + 0.05
    cols = torch.rand(P, 3, generator=g)
    return means, scales, rots, opac, cols
