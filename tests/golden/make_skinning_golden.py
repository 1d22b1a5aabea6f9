#!/usr/bin/env python
"""Generates tests/golden/skinning_*.npz by EXECUTING the reference's own Python for the skinning /
surface-bound-Gaussian path (read from /root/reference at generation time; never at test time).

The reference modules cannot be imported whole (pypose, pytorch3d, open3d, threestudio ... are not
installed — SURVEY.md §8c), so this script
  1. installs small stand-in modules for the third-party ops the path calls (pypose SO3 / so3
     LieTensors, pytorch3d Meshes.faces_normals_list / matrix_to_quaternion) implementing their
     published semantics as restated in SURVEY.md Appendix B.1-B.4,
  2. executes custom/threestudio-dreammesh4d/utils/dual_quaternions.py unmodified,
  3. extracts, by AST, the unmodified source of
        dynamic_sugar.py: strain_tensor_to_matrix, fuse_rotations,
                          DynamicSuGaRModel._get_timed_vertex_attributes_from_dg,
                          .get_timed_vertex_attributes, .get_timed_gs_attributes,
                          ._get_gs_xyz_from_vertex, .get_timed_gs_all_single_time
        sugar.py:         SuGaRModel.points / scaling / quaternions / strengths / get_face_normals /
                          get_gs_normals / get_points_rgb / surface_mesh (+ the trivial getters)
     and runs them on a small seeded mesh + deformation graph,
  4. stores inputs and outputs as fp64 and fp32 fixtures.
What this pins: the reference's own composition of those ops (order, conventions, quirks such as
LBS on the world-space rest vertex, xyzw<->wxyz shuffles, double normalisation).  What it cannot
pin: the third-party ops themselves ("parity unpinned", oracle/skin_oracle.py header).

    python tests/golden/make_skinning_golden.py          # rewrites the .npz files
"""
from __future__ import annotations

import ast
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

REF = Path("/root/reference/custom/threestudio-dreammesh4d")
OUT = Path(__file__).resolve().parent
ROOT = OUT.parents[1]
sys.path.insert(0, str(ROOT))

EPS = 1e-6


# ------------------------------------------------------------------------------------------------
# stand-ins for pypose
# ------------------------------------------------------------------------------------------------
class LieTensor:
    def __init__(self, t, kind):
        self.t = t.t if isinstance(t, LieTensor) else t
        self.kind = kind

    # plumbing
    def tensor(self): return self.t
    @property
    def shape(self): return self.t.shape
    def __getitem__(self, idx): return LieTensor(self.t[idx], self.kind)
    def norm(self, *a, **k): return self.t.norm(*a, **k)
    def to(self, *a, **k): return LieTensor(self.t.to(*a, **k), self.kind)

    # group ops
    def _prod(self, o):
        av, aw, bv, bw = self.t[..., :3], self.t[..., 3:], o.t[..., :3], o.t[..., 3:]
        return LieTensor(torch.cat([aw * bv + bw * av + torch.cross(av, bv, dim=-1),
                                    aw * bw - (av * bv).sum(-1, keepdim=True)], dim=-1), "SO3")

    def __mul__(self, o):
        if isinstance(o, LieTensor): return self._prod(o)
        return self.t * o                     # scalar / tensor: plain tensor arithmetic
    def __rmul__(self, o): return o * self.t
    def __matmul__(self, o): return self._prod(o)
    def __truediv__(self, o): return self.t / (o.t if isinstance(o, LieTensor) else o)
    def __neg__(self): return -self.t
    def Inv(self): return LieTensor(torch.cat([-self.t[..., :3], self.t[..., 3:]], dim=-1), "SO3")

    def Act(self, p):
        v, w = self.t[..., :3], self.t[..., 3:]
        uv = torch.cross(v.expand(*torch.broadcast_shapes(v.shape, p.shape)), p.expand(*torch.broadcast_shapes(v.shape, p.shape)), dim=-1)
        return p + 2 * (w * uv + torch.cross(v.expand_as(uv), uv, dim=-1))

    def matrix(self):
        I = torch.eye(3, dtype=self.t.dtype)
        cols = [self.Act(I[i].expand(*self.t.shape[:-1], 3)) for i in range(3)]
        return torch.stack(cols, dim=-1)

    def Log(self):
        v, w = self.t[..., :3], self.t[..., 3:]
        n = v.norm(dim=-1, keepdim=True)
        big = n > EPS
        ns = torch.where(big, n, torch.ones_like(n))
        f = torch.where(big, 2 * torch.atan(ns / w) / ns, 2 / w - (2.0 / 3.0) * n * n / (w * w * w))
        return LieTensor(f * v, "so3")

    def Exp(self):
        x = self.t
        th = x.norm(dim=-1, keepdim=True)
        big = th > EPS
        ths = torch.where(big, th, torch.ones_like(th))
        th2 = th * th
        a = torch.where(big, torch.sin(0.5 * ths) / ths, 0.5 - th2 / 48 + th2 * th2 / 3840)
        w = torch.where(big, torch.cos(0.5 * ths), 1 - th2 / 8 + th2 * th2 / 384)
        return LieTensor(torch.cat([a * x, w], dim=-1), "SO3")


def install_stubs():
    pp = types.ModuleType("pypose")
    pp.LieTensor = LieTensor
    pp.SO3 = lambda t: LieTensor(t, "SO3")
    pp.so3 = lambda t: LieTensor(t, "so3")
    pp.identity_SO3 = lambda *s: LieTensor(torch.cat([torch.zeros(*s, 3), torch.ones(*s, 1)], -1), "SO3")
    lt = types.ModuleType("pypose.lietensor")
    ltl = types.ModuleType("pypose.lietensor.lietensor")
    ltl.LieType = object
    ltl.SO3Type = object
    pq = types.ModuleType("pyquaternion")
    pq.Quaternion = LieTensor
    sys.modules.update({"pypose": pp, "pypose.lietensor": lt, "pypose.lietensor.lietensor": ltl, "pyquaternion": pq})
    return pp


# ---- stand-ins for pytorch3d ----
class Meshes:
    def __init__(self, verts, faces, textures=None):
        self.v = verts if isinstance(verts, (list, tuple)) else list(verts)
        self.f = faces if isinstance(faces, (list, tuple)) else list(faces)

    def _fn(self, v, f):
        fv = v[f]
        n = torch.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0], dim=-1)
        return n / n.norm(dim=-1, keepdim=True).clamp_min(1e-6)

    def faces_normals_list(self): return [self._fn(v, f) for v, f in zip(self.v, self.f)]
    def faces_normals_padded(self): return torch.stack(self.faces_normals_list(), dim=0)


def matrix_to_quaternion(R):
    from oracle.skin_oracle import matrix_to_quaternion as m2q   # same restated table (Appendix B.4)
    return m2q(R)


class _Ann:
    def __getitem__(self, k): return self
    def __call__(self, *a, **k): return self


def extract(path: Path, names: dict):
    """Returns {name: source} for top-level functions and `Class.method` entries."""
    src = path.read_text()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names.get(None, ()):
            out[node.name] = ast.get_source_segment(src, node)
        if isinstance(node, ast.ClassDef) and node.name in names:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in names[node.name]:
                    seg = ast.get_source_segment(src, sub, padded=True)
                    decos = "".join(ast.get_source_segment(src, d, padded=True).replace(d.id if isinstance(d, ast.Name) else "", "", 0) for d in [])
                    import textwrap
                    body = textwrap.dedent(seg)
                    is_prop = any(isinstance(d, ast.Name) and d.id == "property" for d in sub.decorator_list)
                    out[f"{node.name}.{sub.name}"] = ("@property\n" if is_prop else "") + body
    return out


def main():
    pp = install_stubs()
    ns = {"torch": torch, "F": F, "pp": pp, "nn": torch.nn, "np": np, "Float": _Ann(), "Int": _Ann(), "Tensor": _Ann(),
          "Dict": _Ann(), "Union": _Ann(), "Any": _Ann(), "Meshes": Meshes, "TexturesVertex": lambda **k: None,
          "matrix_to_quaternion": matrix_to_quaternion, "__name__": "ref_exec"}
    from einops import rearrange
    ns["rearrange"] = rearrange
    # (2) dual quaternions, unmodified
    exec(compile((REF / "utils/dual_quaternions.py").read_text(), "dual_quaternions.py", "exec"), ns)
    # SH2RGB (geometry/gaussian_base.py:32-40)
    gb = extract(REF / "geometry/gaussian_base.py", {None: ("SH2RGB", "RGB2SH")})
    ns["C0"] = 0.28209479177387814
    for s in gb.values():
        exec(s, ns)
    # (3) reference functions / methods, unmodified source
    dyn = extract(REF / "geometry/dynamic_sugar.py", {
        None: ("strain_tensor_to_matrix", "fuse_rotations", "dict_temporal_key"),
        "DynamicSuGaRModel": ("_get_timed_vertex_attributes_from_dg", "get_timed_vertex_attributes",
                              "get_timed_gs_attributes", "_get_gs_xyz_from_vertex", "get_timed_gs_all_single_time",
                              "get_timed_vertex_xyz", "get_timed_surface_mesh", "get_timed_face_normals",
                              "get_timed_gs_normals")})
    sug = extract(REF / "geometry/sugar.py", {"SuGaRModel": (
        "points", "scaling", "quaternions", "strengths", "get_face_normals", "get_gs_normals", "get_points_rgb",
        "surface_mesh", "get_scaling", "get_opacity", "get_rotation", "get_xyz", "get_xyz_verts", "n_faces", "n_verts")})
    for k in ("strain_tensor_to_matrix", "fuse_rotations", "dict_temporal_key"):
        exec(dyn[k], ns)

    class Ref:      # bare host object: only the attributes the extracted methods touch
        pass
    for table in (sug, dyn):
        for k, s in table.items():
            if "." in k:
                loc = {}
                exec(s, ns, loc)
                setattr(Ref, k.split(".")[1], list(loc.values())[0])

    from dreammesh4d_b200 import synthetic
    for tag, dtype, g, method, d_scale in (("f64_g3_hybrid", torch.float64, 3, "hybrid", False), ("f32_g6_hybrid", torch.float32, 6, "hybrid", False),
                                           ("f64_g3_lbs", torch.float64, 3, "lbs", False), ("f64_g3_dqs", torch.float64, 3, "dqs", False),
                                           ("f64_g3_hybrid_dscale", torch.float64, 3, "hybrid", True),
                                           ("f64_g3_lbs_dscale", torch.float64, 3, "lbs", True)):
        torch.manual_seed(0)
        scene = synthetic.make_sugar_scene(264, g=g)        # 12 x 11 UV sphere
        gen = torch.Generator().manual_seed(3)
        scene.verts = scene.verts + 0.01 * torch.randn(scene.verts.shape, generator=gen)
        scene.complex_rot = F.normalize(torch.randn(scene.complex_rot.shape, generator=gen), dim=-1)
        scene.log_scales = scene.log_scales + 0.3 * torch.randn(scene.log_scales.shape, generator=gen)
        scene.densities = torch.randn(scene.densities.shape, generator=gen)
        graph = synthetic.make_deform_graph(scene.verts, 24, 4, seed=0)
        T = 2
        trans, rot, scale, opac = synthetic.random_node_attrs(T, 24, seed=1)
        rot = F.normalize(rot + 0.3 * torch.randn(rot.shape, generator=gen), dim=-1)    # larger rotations
        c = lambda t: t.to(dtype)

        r = Ref()
        r.cfg = types.SimpleNamespace(skinning_method=method, d_scale=d_scale, use_deform_graph=True,
                                      n_gaussians_per_surface_triangle=g, sh_levels=1)
        r.device = "cpu"
        r.binded_to_surface_mesh = True
        r._points = c(scene.verts)
        r._surface_mesh_faces = scene.faces
        r.surface_triangle_bary_coords = c(scene.bary)[..., None]
        r._n_points = scene.n_gaussians
        r._scales = c(scene.log_scales)
        r._quaternions = c(scene.complex_rot)
        r.all_densities = c(scene.densities)
        r._sh_coordinates_dc = c(scene.sh_dc)
        r._vertex_colors = torch.zeros_like(r._points)
        r.surface_mesh_thickness = torch.tensor(scene.thickness, dtype=dtype)
        r.scale_activation = torch.exp
        r._deform_graph_node_xyz = c(graph.node_xyz)
        r._xyz_neighbor_node_idx = graph.nbr_idx
        r._xyz_neighbor_nodes_weights = c(graph.nbr_w)
        r._gs_bary_weights = torch.cat([r.surface_triangle_bary_coords] * scene.faces.shape[0], dim=0)     # dynamic_sugar.py:154-156
        r._gs_vert_connections = scene.faces.repeat_interleave(g, dim=0)                                    # :157-159
        r._deformed_vert_positions, r._deformed_vert_rotations = {}, {}
        attrs = {"xyz": c(trans), "rotation": pp.SO3(c(rot)), "scale": c(scale), "opacity": c(opac)}
        r.get_timed_dg_attributes = lambda timestamp, frame_idx: {k: (v[:len(timestamp)] if not isinstance(v, LieTensor) else v[:len(timestamp)]) for k, v in attrs.items()}

        ts = torch.linspace(0, 1, T + 2)[1:-1].to(dtype)
        vert = r._get_timed_vertex_attributes_from_dg(ts, None)
        gs = r.get_timed_gs_attributes(ts, None)
        normals = r.get_timed_gs_normals(ts, None)
        # single-time path exactly as the renderer calls it (diff_sugar_rasterizer_temporal.py:161)
        attrs1 = {k: v[1:2] for k, v in attrs.items()}
        r.get_timed_dg_attributes = lambda timestamp, frame_idx: attrs1
        m1, s1, r1, o1, c1 = r.get_timed_gs_all_single_time(ts[1], None)
        np.savez_compressed(
            OUT / f"skinning_{tag}.npz",
            # inputs
            verts=r._points.numpy(), faces=scene.faces.numpy(), bary=c(scene.bary).numpy(), log_scales=r._scales.numpy(),
            complex_rot=r._quaternions.numpy(), densities=r.all_densities.numpy(), sh_dc=r._sh_coordinates_dc.numpy(),
            thickness=np.asarray(scene.thickness), g=np.asarray(g), method=np.asarray(method),
            nbr_idx=graph.nbr_idx.numpy(), nbr_w=r._xyz_neighbor_nodes_weights.numpy(),
            node_trans=c(trans).numpy(), node_rot=c(rot).numpy(), node_scale=c(scale).numpy(), node_opacity=c(opac).numpy(),
            # outputs of the reference code
            out_vert_xyz=vert["xyz"].numpy(), out_vert_rot=vert["rotation"].tensor().numpy(),
            out_gs_xyz=gs["xyz"].numpy(), out_gs_rot=gs["rotation"].numpy(), out_gs_normals=normals.numpy(),
            out_static_xyz=r.get_xyz.numpy(), out_static_scaling=r.get_scaling.numpy(), out_static_rot=r.get_rotation.numpy(),
            out_static_opacity=r.get_opacity.numpy(), out_static_rgb=r.get_points_rgb().numpy(),
            out_static_normals=r.get_gs_normals.numpy(),
            out_single_means=m1.numpy(), out_single_scales=s1.numpy(), out_single_rot=r1.numpy(),
            out_single_opacity=o1.numpy(), out_single_colors=c1.numpy(),
            **({"out_vert_scale": vert["scale"].numpy(), "out_gs_scale": gs["scale"].numpy()} if d_scale else {}),
        )
        print("wrote", f"skinning_{tag}.npz", "V", r._points.shape[0], "P", scene.n_gaussians)

    # strain_tensor_to_matrix (dynamic_sugar.py:29-39)
    st = torch.randn(2, 5, 6, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    np.savez_compressed(OUT / "strain.npz", strain=st.numpy(), matrix=ns["strain_tensor_to_matrix"](st).numpy())
    print("wrote strain.npz")


if __name__ == "__main__":
    main()
