"""Mesh ingest for the surface-bound Gaussians (start-up path of SuGaRModel.configure,
custom/threestudio-dreammesh4d/geometry/sugar.py:74-117,119-161,166-233,300-325): read a triangle mesh, keep its
dominant connected component, bind ``g`` Gaussians to every face.

The reference reads with Open3D and prunes with a Python BFS restarted from every vertex (O(V^2) worst case); here the
component labelling is one union-find / sparse-graph call.  Open3D is used when it is installed (same reader as the
reference); otherwise ASCII OBJ / PLY files are parsed directly — enough for the meshes the static stage exports.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

from .synthetic import BARY, C0, SugarScene, circle_radius


def read_triangle_mesh(path) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(vertices [V,3] float64, triangles [F,3] int64, vertex colours [V,3] in [0,1] or an empty array)."""
    if not isinstance(path, (str, Path)):                      # an Open3D-like object (sugar.py:74 passes one through)
        return (np.asarray(path.vertices, dtype=np.float64), np.asarray(path.triangles, dtype=np.int64),
                np.asarray(getattr(path, "vertex_colors", np.zeros((0, 3))), dtype=np.float64))
    try:
        import open3d as o3d
        m = o3d.io.read_triangle_mesh(str(path))
        return np.asarray(m.vertices), np.asarray(m.triangles).astype(np.int64), np.asarray(m.vertex_colors)
    except ImportError:
        pass
    p = Path(path)
    if p.suffix.lower() == ".obj":
        v, c, f = [], [], []
        for line in p.read_text().splitlines():
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                v.append([float(x) for x in t[1:4]])
                if len(t) >= 7:
                    c.append([float(x) for x in t[4:7]])
            elif t[0] == "f":
                idx = [int(x.split("/")[0]) for x in t[1:]]
                idx = [i - 1 if i > 0 else len(v) + i for i in idx]
                for k in range(1, len(idx) - 1):               # fan-triangulate polygons
                    f.append([idx[0], idx[k], idx[k + 1]])
        cols = np.asarray(c, dtype=np.float64) if len(c) == len(v) else np.zeros((0, 3))
        return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int64), cols
    if p.suffix.lower() == ".ply":
        lines = p.read_text(errors="ignore").splitlines()
        if "format ascii" not in "\n".join(lines[:5]):
            raise ValueError("binary PLY needs open3d; export ASCII PLY or OBJ")
        nv = nf = 0
        props, in_vertex, end = [], False, 0
        for i, line in enumerate(lines):
            t = line.split()
            if t[:2] == ["element", "vertex"]:
                nv, in_vertex = int(t[2]), True
            elif t[:2] == ["element", "face"]:
                nf, in_vertex = int(t[2]), False
            elif t and t[0] == "property" and in_vertex:
                props.append(t[-1])
            elif t and t[0] == "end_header":
                end = i + 1
                break
        vals = np.array([[float(x) for x in l.split()] for l in lines[end:end + nv]], dtype=np.float64)
        col = {n: k for k, n in enumerate(props)}
        verts = vals[:, [col["x"], col["y"], col["z"]]]
        cols = vals[:, [col["red"], col["green"], col["blue"]]] / 255.0 if "red" in col else np.zeros((0, 3))
        faces = []
        for l in lines[end + nv:end + nv + nf]:
            t = [int(x) for x in l.split()]
            for k in range(2, t[0]):
                faces.append([t[1], t[k], t[k + 1]])
        return verts, np.asarray(faces, dtype=np.int64), cols
    raise ValueError(f"unsupported mesh file {p.name}: install open3d or use .obj / ASCII .ply")


def keep_dominant_component(verts: np.ndarray, faces: np.ndarray, colors: np.ndarray):
    """sugar.py:119-161 — keep the first connected component (in vertex order) holding more than ceil(0.75 V) vertices;
    vertices re-indexed in ascending order, faces that lose a vertex dropped.  If no component is that large the
    reference ends up with the component of the LAST vertex; so does this."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    V = len(verts)
    i = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2]])
    j = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0]])
    _, label = connected_components(coo_matrix((np.ones(len(i)), (i, j)), shape=(V, V)), directed=False)
    counts = np.bincount(label)
    big = np.flatnonzero(counts > math.ceil(V * 0.75))
    chosen = big[0] if len(big) else label[V - 1]
    keep = np.flatnonzero(label == chosen)
    remap = -np.ones(V, dtype=np.int64)
    remap[keep] = np.arange(len(keep))
    f = remap[faces]
    f = f[(f >= 0).all(axis=1)]
    return verts[keep], f, (colors[keep] if len(colors) == V else colors)


def scene_from_mesh(verts, faces, colors, g: int, init_gs_scales_s: float = 1.7, init_gs_opacity: float = 0.5,
                    spatial_extent: float = 3.5, inherit_vertex_colors: bool = True,
                    learn_opacity: bool = True) -> SugarScene:
    """The static SuGaR state the reference builds in load_surface_mesh_to_bind / configure / initialize_learnable_radiuses
    (sugar.py:166-233, 96-106, 300-325) for a mesh: barycentric table, SH DC from interpolated vertex colours, densities
    = logit(init opacity) (logit(0.9999) when opacities are frozen), log radii from the shortest edge, unit complex rotations."""
    verts = torch.as_tensor(np.asarray(verts), dtype=torch.float32)
    faces = torch.as_tensor(np.asarray(faces), dtype=torch.long)
    if len(colors) != len(verts):
        colors = np.ones((len(verts), 3)) * 0.5                                   # sugar.py:183-184
    vcol = torch.as_tensor(np.asarray(colors), dtype=torch.float32)
    bary = torch.tensor(BARY[g], dtype=torch.float32)
    P = faces.shape[0] * g
    if inherit_vertex_colors:
        col = (vcol[faces][:, None] * bary[None, :, :, None]).sum(dim=-2).reshape(-1, 3)
    else:
        col = torch.full((P, 3), 0.5)
    sh_dc = ((col - 0.5) / C0)[:, None, :].contiguous()
    fv = verts[faces]
    edge = (fv - fv[:, [1, 2, 0]]).norm(dim=-1).min(dim=-1)[0]
    scales = (edge * circle_radius(g, init_gs_scales_s)).clamp_min(1e-7)
    log_scales = scales.log()[:, None, None].expand(-1, g, 2).reshape(-1, 2).contiguous()
    complex_rot = torch.zeros(P, 2)
    complex_rot[:, 0] = 1.0
    op = init_gs_opacity if learn_opacity else 0.9999
    densities = torch.full((P, 1), math.log(op / (1 - op)))
    scene = SugarScene(verts, faces, bary, log_scales, complex_rot, densities, sh_dc, spatial_extent / 1_000_000, g)
    scene.vertex_colors = vcol
    return scene


def load_scene(path_or_mesh, g: int, **kw) -> SugarScene:
    v, f, c = read_triangle_mesh(path_or_mesh)
    v, f, c = keep_dominant_component(v, f, c)
    return scene_from_mesh(v, f, c, g, **kw)
