"""B-outer boundary: the registered threestudio classes of dreammesh4d_b200/plugin.py against the
reference's own class definitions and the call sites of its systems (tests/golden/plugin_api.json, extracted from
/root/reference by AST).  threestudio is not installed here: the classes run on the stand-in base with the same
constructor protocol (``cls(cfg_dict, geometry=...)`` -> ``configure``)."""
import dataclasses
import json
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from dreammesh4d_b200 import plugin, synthetic

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "plugin_api.json").read_text())
# attributes of the free-Gaussian ("gaussian" stage) branches of the static system, dead at the shipped configs
OUT_OF_SCOPE = {"create_from_pcd", "save_ply", "update_states"}


@pytest.mark.parametrize("name", sorted(GOLD["classes"]))
def test_registered_names_and_config_fields_match_the_reference(name):
    ref = GOLD["classes"][name]
    cls = plugin.REGISTRY[name]
    assert cls.__name__ == ref["class"]
    own = {f.name: f for f in dataclasses.fields(cls.Config)}
    for fname, default in ref["config"]:
        assert fname in own, f"{name}: Config.{fname} missing"
        d = own[fname].default
        d = list(d) if isinstance(d, tuple) else d
        assert d == default, f"{name}: Config.{fname} default {d!r} != reference {default!r}"


def test_unknown_config_keys_raise_like_structured_configs():
    if plugin.HAVE_THREESTUDIO:
        pytest.skip("threestudio's own OmegaConf parsing applies")
    with pytest.raises(KeyError):
        plugin.DiffGaussian({"no_such_key": 1}, geometry=None)


@pytest.mark.parametrize("system", sorted(GOLD["system_calls"]))
def test_everything_the_systems_touch_exists(system):
    calls = GOLD["system_calls"][system]
    geo_cls = plugin.DynamicSuGaRModel if "4dgen" in system else plugin.SuGaRModel
    for attr in calls["geometry"]:
        if attr in OUT_OF_SCOPE:
            continue
        assert hasattr(geo_cls, attr) or attr in ("optimizer", "_vertex_colors"), f"geometry.{attr} (read by {system})"
    for attr in calls["renderer"]:
        assert hasattr(plugin.DiffGaussian, attr) and hasattr(plugin.DiffSuGaR, attr)


def _mesh(n_faces=2000):
    verts, faces = synthetic.uv_sphere(n_faces)
    col = 0.5 + 0.5 * torch.nn.functional.normalize(verts, dim=-1)
    # an Open3D-like object, as SuGaRModel.configure(o3d_mesh) accepts (sugar.py:74,175-181) — plus an isolated sliver
    v = np.concatenate([verts.numpy(), np.array([[5, 5, 5], [5.1, 5, 5], [5, 5.1, 5]])])
    f = np.concatenate([faces.numpy(), np.array([[len(verts), len(verts) + 1, len(verts) + 2]])])
    c = np.concatenate([col.numpy(), np.ones((3, 3)) * 0.5])
    return types.SimpleNamespace(vertices=v, triangles=f, vertex_colors=c), len(verts), len(faces)


def test_mesh_ingest_keeps_the_dominant_component_and_binds_gaussians():
    from dreammesh4d_b200 import mesh_io
    mesh, V, F = _mesh()
    scene = mesh_io.load_scene(mesh, 6, init_gs_scales_s=1.3, init_gs_opacity=0.9)
    assert scene.verts.shape[0] == V and scene.faces.shape[0] == F            # the isolated triangle is gone
    assert scene.n_gaussians == 6 * F and scene.sh_dc.shape == (6 * F, 1, 3)
    ref = synthetic.make_sugar_scene(2000, g=6)                                # same sphere through the synthetic builder
    assert torch.allclose(scene.log_scales, ref.log_scales) and torch.allclose(scene.densities, ref.densities)
    assert torch.allclose(scene.sh_dc, ref.sh_dc, atol=1e-6)


@pytest.mark.gpu
def test_dynamic_stage_objects_build_from_cfg_and_train():
    """configs/sugar_dynamic_dg.yaml's geometry / renderer blocks (resolved values), a mesh object instead of a file."""
    from dreammesh4d_b200.trainstep import DynamicStageStep
    mesh, V, F = _mesh()
    geo_cfg = dict(num_frames=8, use_deform_graph=True, dynamic_mode="deformation", n_dg_nodes=64, dg_node_connectivity=4,
                   deformation_lr=0.00064, grid_lr=0.0064, d_xyz=True, d_rotation=True, d_opacity=False, d_scale=False,
                   dist_mode="geodisc", skinning_method="hybrid", position_lr=0.0001, spatial_lr_scale=1.0,
                   spatial_extent=1.0, n_gaussians_per_surface_triangle=3, init_gs_scales_s=1.3, init_gs_opacity=0.9,
                   surface_mesh_to_bind_path="unused-when-a-mesh-object-is-passed.obj")
    geo = plugin.REGISTRY["dynamic-sugar"](geo_cfg, mesh)
    ren = plugin.REGISTRY["diff-sugar-rasterizer-temporal"]({"back_ground_color": (1.0, 1.0, 1.0)}, geometry=geo, material=None, background=None)
    # checkpoint schema + the graph buffers
    keys = set(geo.state_dict().keys())
    for k in ("_surface_mesh_faces", "surface_mesh_thickness", "_points", "_sh_coordinates_dc", "_sh_coordinates_rest",
              "all_densities", "_scales", "_quaternions", "_deformation.deformation_net.grid.grids.0.0",
              "_deformation.deformation_net.pos_deform.feature_out.1.weight", "_deform_graph_node_xyz",
              "_xyz_neighbor_node_idx", "_xyz_neighbor_nodes_weights"):
        assert k in keys, k
    assert not any(p.requires_grad for n, p in geo.named_parameters() if not n.startswith("_deformation"))   # frozen statics
    assert [g["name"] for g in geo.optimizer.param_groups] == ["deformation", "grid"]
    assert geo.get_xyz.shape[0] == 3 * F and geo._xyz_neighbor_node_idx.shape == (V, 4)
    assert torch.allclose(geo._xyz_neighbor_nodes_weights.sum(-1), torch.ones(V, device="cuda"), atol=1e-5)
    lr0 = geo.optimizer.param_groups[1]["lr"]
    geo.cfg.grid_lr = [0, 0.0064, 0.00064, 1000]
    geo.update_learning_rate(500)
    assert geo.optimizer.param_groups[1]["lr"] == pytest.approx(0.0064 * (0.1 ** 0.5)) and lr0 == pytest.approx(0.0064)
    opt = geo.merge_optimizer(torch.optim.Adam([torch.nn.Parameter(torch.zeros(3, device="cuda"))], lr=0.01))
    assert isinstance(opt, torch.optim.AdamW) and len(opt.param_groups) == 3
    # one batch through renderer.batch_forward, as SuGaR4DGen.forward does (sugar_4dgen.py:78-81)
    B, H = 2, 64
    c2w, fovy = synthetic.random_orbit_cameras(B, seed=1)
    with torch.no_grad():
        for head in (geo._deformation.deformation_net.pos_deform, geo._deformation.deformation_net.rotations_deform):
            head.feature_out[1].weight.normal_(0, 0.02)
    batch = {"c2w": c2w.cuda(), "fovy": fovy.cuda(), "height": H, "width": H, "timestamp": torch.tensor([0.25, 0.75], device="cuda"),
             "frame_indices": torch.tensor([1, 5], device="cuda")}
    geo.update_learning_rate(0)
    out = ren.batch_forward(batch)
    assert out["comp_rgb"].shape == (B, H, H, 3) and out["comp_mask"].shape == (B, H, H, 1) and len(out["radii"]) == B
    assert float(out["comp_mask"].max()) > 0.5
    meshes = geo.get_timed_surface_mesh(timestamp=batch["timestamp"], frame_idx=batch["frame_indices"])
    assert meshes.verts_padded().shape == (B, V, 3)
    assert geo.get_timed_vertex_rotation(batch["timestamp"], batch["frame_indices"], return_matrix=True).shape == (B, V, 3, 3)
    assert len(geo._deformed_vert_positions) == B
    # and a full optimizer step moves the deformation network only
    before = [p.detach().clone() for p in geo._deformation.parameters()]
    target = torch.rand(B, H, H, 3, device="cuda")
    step = DynamicStageStep(geo, ren, opt, lambda o, b: torch.nn.functional.mse_loss(o["comp_rgb"], target) * 100)
    step([batch], 0)
    moved = sum(float((a - b).abs().max()) > 0 for a, b in zip(geo._deformation.parameters(), before))
    assert moved >= 6
    geo.update_step(0, 1)
    assert geo._timed is None


@pytest.mark.gpu
def test_static_stage_objects_build_from_cfg():
    mesh, V, F = _mesh()
    geo = plugin.REGISTRY["sugar"](dict(n_gaussians_per_surface_triangle=6, spatial_extent=1.0, init_gs_opacity=0.9,
                                        surface_mesh_to_bind_path="x.obj"), mesh)
    ren = plugin.REGISTRY["diff-sugar-rasterizer-normal"]({}, geometry=geo, material=None, background=None)
    assert [g["name"] for g in geo.optimizer.param_groups] == ["points", "f_dc", "f_rest", "all_densities", "scales", "quaternions"]
    c2w, fovy = synthetic.random_orbit_cameras(2, seed=3)
    out = ren.batch_forward({"c2w": c2w.cuda(), "fovy": fovy.cuda(), "height": 64, "width": 64})
    (out["comp_rgb"].mean() + out["comp_normal"].mean()).backward()
    assert geo._points.grad is not None and float(geo._points.grad.abs().max()) > 0
    assert geo.surface_mesh.verts_list()[0].shape == (V, 3)


@pytest.mark.gpu
def test_d_scale_true_flows_through_the_batched_renderer():
    """Class default ``d_scale=True`` (dynamic_sugar.py:68): per-timestamp Gaussian scales reach the rasterizer as one
    attribute set per view and their gradient reaches the scale head of the deformation network."""
    mesh, V, F = _mesh()
    geo = plugin.REGISTRY["dynamic-sugar"](dict(n_dg_nodes=32, dg_node_connectivity=4, d_scale=True, skinning_method="hybrid",
                                               spatial_extent=1.0, n_gaussians_per_surface_triangle=3, init_gs_opacity=0.9,
                                               surface_mesh_to_bind_path="x.obj"), mesh)
    ren = plugin.REGISTRY["diff-sugar-rasterizer-temporal"]({}, geometry=geo, material=None, background=None)
    with torch.no_grad():
        geo._deformation.deformation_net.scales_deform.feature_out[1].weight.normal_(0, 0.05)
    c2w, fovy = synthetic.random_orbit_cameras(2, seed=5)
    batch = {"c2w": c2w.cuda(), "fovy": fovy.cuda(), "height": 64, "width": 64, "timestamp": torch.tensor([0.3, 0.6], device="cuda")}
    out = ren.batch_forward(batch)
    assert geo._timed["scales"].shape == (2, 3 * F, 3)
    assert not torch.allclose(geo._timed["scales"][0], geo.get_scaling)
    out["comp_rgb"].square().mean().backward()
    g = geo._deformation.deformation_net.scales_deform.feature_out[1].weight.grad
    assert g is not None and float(g.abs().max()) > 0
