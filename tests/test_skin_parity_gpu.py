"""GPU parity: fused skinning + surface-bound Gaussian update (libdm4d.so) vs the torch oracle (fp64),
on the golden-vector inputs (which pin the oracle to the reference's own code) and on a larger mesh.
Tolerances: 1e-4 relative L-inf forward (fp32 kernel vs fp64 oracle), 1e-3 on gradients (north-star)."""
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from dreammesh4d_b200 import skinning, synthetic
from oracle import skin_oracle as SO
from tests import helpers as Hh

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = Path(__file__).resolve().parent / "golden"
TOL_FWD = 1e-4


def qalign(a, b):
    return a * torch.sign((a * b).sum(-1, keepdim=True))


def run_both(scene, graph, node, method, seed=0, with_vertex_grads=True):
    trans, rot, scale, opac = [t.double().requires_grad_(True) for t in node]
    out = SO.deform_gaussians(scene, graph, trans, rot, scale, opac, method=method, dtype=torch.float64)
    g = torch.Generator().manual_seed(seed)
    T, P, V = out["means3D"].shape[0], out["means3D"].shape[1], out["verts"].shape[1]
    gm, gr, gn = (torch.randn(T, P, k, generator=g, dtype=torch.float64) for k in (3, 4, 3))
    # fp32 conditioning of a face normal: an absolute vertex error d moves the unit normal by about
    # d * (|e1| + |e2|) / |e1 x e2| (sliver triangles at the poles of the UV sphere amplify it); the
    # normal gradient is only probed on well-conditioned faces and the forward bound is per face.
    fv = out["verts"].detach()[:, scene.faces]
    e1, e2 = fv[:, :, 1] - fv[:, :, 0], fv[:, :, 2] - fv[:, :, 0]
    cond = (e1.norm(dim=-1) + e2.norm(dim=-1)) / torch.cross(e1, e2, dim=-1).norm(dim=-1).clamp_min(1e-30)
    nbound = (4e-7 * cond).repeat_interleave(scene.g, dim=1)[..., None]          # [T,P,1]
    gn = gn * (nbound < 2e-5)
    gv, gvr = (torch.randn(T, V, k, generator=g, dtype=torch.float64) * (1.0 if with_vertex_grads else 0.0) for k in (3, 4))
    loss = (out["means3D"] * gm).sum() + (out["rotations"] * gr).sum() + (out["normals"] * gn).sum() + \
        (out["verts"] * gv).sum() + (out["vert_rot"] * gvr).sum()
    loss.backward()

    d = lambda t: t.to(DEV)
    rq, nrm0 = skinning.sugar_rest_frames(d(scene.verts), d(scene.faces.int()), d(scene.complex_rot), scene.g)
    assert Hh.rel_linf(qalign(rq.cpu().double(), out["rest_quat"].detach()), out["rest_quat"].detach()) <= TOL_FWD
    ct = [d(t.detach().float()).requires_grad_(True) for t in (trans, rot, scale, opac)]
    means, rots, normals, verts, vrot = skinning.skin_gaussians(
        *ct, d(scene.verts), d(scene.faces.int()), d(graph.nbr_idx.int()), d(graph.nbr_w), d(scene.bary),
        d(out["rest_quat"].detach().float()), method=method)
    for name, got, ref in (("verts", verts, out["verts"]), ("means3D", means, out["means3D"])):
        assert Hh.rel_linf(got.detach().cpu().double(), ref.detach()) <= TOL_FWD, name
    nerr = (normals.detach().cpu().double() - out["normals"].detach()).abs()
    assert bool((nerr <= TOL_FWD + nbound).all()), f"normals: worst excess {(nerr - nbound).max().item()}"
    for name, got, ref in (("vert_rot", vrot, out["vert_rot"]), ("rotations", rots, out["rotations"])):
        assert Hh.rel_linf(qalign(got.detach().cpu().double(), ref.detach()), ref.detach()) <= TOL_FWD, name
    # the kernel and the oracle may differ by the (irrelevant) quaternion sign; feed gradients consistently
    s_r = torch.sign((rots.detach().cpu().double() * out["rotations"].detach()).sum(-1, keepdim=True))
    s_v = torch.sign((vrot.detach().cpu().double() * out["vert_rot"].detach()).sum(-1, keepdim=True))
    assert (s_r > 0).all() and (s_v > 0).all()
    loss2 = (means * d(gm.float())).sum() + (rots * d(gr.float())).sum() + (normals * d(gn.float())).sum() + \
        (verts * d(gv.float())).sum() + (vrot * d(gvr.float())).sum()
    loss2.backward()
    names = ["node_trans", "node_rot", "node_scale", "node_opacity"]
    for name, got, ref in zip(names, ct, (trans, rot, scale, opac)):
        if method == "dqs" and name in ("node_scale", "node_opacity"):
            assert float(got.grad.abs().max()) == 0.0
            continue
        if method == "lbs" and name == "node_opacity":
            assert float(got.grad.abs().max()) == 0.0
            continue
        err = Hh.rel_linf(got.grad.cpu().double(), ref.grad)
        assert err <= Hh.TOL_GRAD, f"{method} grad {name}: rel Linf {err}"


@pytest.mark.parametrize("tag", ["f64_g3_hybrid", "f32_g6_hybrid", "f64_g3_lbs", "f64_g3_dqs"])
def test_skin_on_golden_inputs(tag):
    z = np.load(GOLD / f"skinning_{tag}.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files if k != "method"}
    g = int(z["g"])
    scene = types.SimpleNamespace(verts=t["verts"].float(), faces=t["faces"], bary=t["bary"].float(),
                                  log_scales=t["log_scales"].float(), complex_rot=t["complex_rot"].float(),
                                  densities=t["densities"].float(), sh_dc=t["sh_dc"].float(),
                                  thickness=float(z["thickness"]), g=g)
    graph = types.SimpleNamespace(nbr_idx=t["nbr_idx"], nbr_w=t["nbr_w"].float())
    node = (t["node_trans"].float(), t["node_rot"].float(), t["node_scale"].float(), t["node_opacity"].float())
    run_both(scene, graph, node, str(z["method"]))
    # and directly against the reference-code outputs stored in the fixture
    d = lambda x: x.to(DEV)
    rq, _ = skinning.sugar_rest_frames(d(scene.verts), d(scene.faces.int()), d(scene.complex_rot), g)
    means, rots, normals, verts, vrot = skinning.skin_gaussians(
        *[d(x) for x in node], d(scene.verts), d(scene.faces.int()), d(graph.nbr_idx.int()), d(graph.nbr_w),
        d(scene.bary), rq, method=str(z["method"]))
    assert Hh.rel_linf(verts.cpu().double(), t["out_vert_xyz"].double()) <= TOL_FWD
    assert Hh.rel_linf(means.cpu().double(), t["out_gs_xyz"].double()) <= TOL_FWD
    assert Hh.rel_linf(normals.cpu().double(), t["out_gs_normals"].double()) <= TOL_FWD
    assert Hh.rel_linf(qalign(rots.cpu().double(), t["out_gs_rot"].double()), t["out_gs_rot"].double()) <= TOL_FWD
    assert Hh.rel_linf(qalign(rq.cpu().double(), t["out_static_rot"].double()), t["out_static_rot"].double()) <= TOL_FWD


@pytest.mark.parametrize("method", ["hybrid", "lbs", "dqs"])
def test_skin_larger_mesh(method):
    scene = synthetic.make_sugar_scene(20_000, g=3)
    graph = synthetic.make_deform_graph(scene.verts, 128, 4)
    node = synthetic.random_node_attrs(3, 128, seed=4)
    run_both(scene, graph, node, method, seed=2)


def test_identity_deformation_is_identity_gpu():
    scene = synthetic.make_sugar_scene(2_000, g=6)
    graph = synthetic.make_deform_graph(scene.verts, 32, 8)
    T, M = 2, 32
    d = lambda x: x.to(DEV)
    rot = torch.zeros(T, M, 4); rot[..., 3] = 1
    rq, _ = skinning.sugar_rest_frames(d(scene.verts), d(scene.faces.int()), d(scene.complex_rot), 6)
    means, rots, normals, verts, vrot = skinning.skin_gaussians(
        d(torch.zeros(T, M, 3)), d(rot), d(torch.eye(3).expand(T, M, 3, 3).contiguous()), d(torch.full((T, M, 1), 0.3)),
        d(scene.verts), d(scene.faces.int()), d(graph.nbr_idx.int()), d(graph.nbr_w), d(scene.bary), rq)
    assert (verts.cpu() - scene.verts[None]).abs().max() < 1e-6
    assert (qalign(rots, rq[None]) - rq[None]).abs().max() < 1e-6


def test_rest_frames_backward_static_stage():
    """Static stage (BASELINE config 1): vertices and in-plane rotations are learnable; gradients of the
    rest-pose quaternions / normals vs oracle autograd (fp64)."""
    scene = synthetic.make_sugar_scene(1_200, g=3)
    gen = torch.Generator().manual_seed(5)
    verts = (scene.verts + 0.01 * torch.randn(scene.verts.shape, generator=gen)).double().requires_grad_(True)
    cplx = torch.randn(scene.complex_rot.shape, generator=gen).double().requires_grad_(True)
    q_ref = SO.sugar_quaternions(verts, scene.faces, cplx, scene.g)
    n_ref = SO.faces_normals(verts, scene.faces).repeat_interleave(scene.g, dim=0)
    gq = torch.randn(q_ref.shape, generator=gen).double()
    gn = torch.randn(n_ref.shape, generator=gen).double()
    ((q_ref * gq).sum() + (n_ref * gn).sum()).backward()

    v = verts.detach().float().to(DEV).requires_grad_(True)
    c = cplx.detach().float().to(DEV).requires_grad_(True)
    q, n = skinning.sugar_rest_frames(v, scene.faces.int().to(DEV), c, scene.g)
    sgn = torch.sign((q.detach().cpu().double() * q_ref.detach()).sum(-1, keepdim=True))
    assert Hh.rel_linf(q.detach().cpu().double() * sgn, q_ref.detach()) <= TOL_FWD
    assert Hh.rel_linf(n.detach().cpu().double(), n_ref.detach()) <= TOL_FWD
    ((q * (gq * sgn).float().to(DEV)).sum() + (n * gn.float().to(DEV)).sum()).backward()
    assert Hh.rel_linf(v.grad.cpu().double(), verts.grad) <= Hh.TOL_GRAD
    assert Hh.rel_linf(c.grad.cpu().double(), cplx.grad) <= Hh.TOL_GRAD


def test_node_incidence_lists_and_list_free_backward():
    """dm4d_skin_node_incidence vs a numpy stable sort (both regimes); the gather-based backward (incidence lists) is
    bit-reproducible and agrees with the list-free backward (the C ABI's path when the desc carries no lists)."""
    scene = synthetic.make_sugar_scene(6_000, g=3)
    M, T = 160, 8
    graph = synthetic.make_deform_graph(scene.verts, M, 4)
    nbr = graph.nbr_idx.int().to(DEV).contiguous()
    inc_ptr, inc = skinning.node_incidence(nbr, M)
    flat = graph.nbr_idx.numpy().reshape(-1)
    order = np.argsort(flat, kind="stable").astype(np.int32)
    counts = np.bincount(flat, minlength=M)
    assert np.array_equal(inc.cpu().numpy(), order)
    assert np.array_equal(inc_ptr.cpu().numpy(), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))

    d = lambda x: x.to(DEV)
    node = synthetic.random_node_attrs(T, M, seed=9)
    rq, _ = skinning.sugar_rest_frames(d(scene.verts), d(scene.faces.int()), d(scene.complex_rot), scene.g)
    gen = torch.Generator().manual_seed(3)
    P = scene.faces.shape[0] * scene.g
    gm, gr, gn = (torch.randn(T, P, k, generator=gen).to(DEV) for k in (3, 4, 3))

    def grads(use_lists):
        skinning.USE_NODE_INCIDENCE = use_lists
        skinning.REPRODUCIBLE = use_lists
        try:
            ct = [d(t.float()).requires_grad_(True) for t in node]
            means, rots, normals, _, _ = skinning.skin_gaussians(*ct, d(scene.verts), d(scene.faces.int()), nbr, d(graph.nbr_w),
                                                                 d(scene.bary), rq, method="hybrid")
            ((means * gm).sum() + (rots * gr).sum() + (normals * gn).sum()).backward()
            return [c.grad.clone() for c in ct]
        finally:
            skinning.USE_NODE_INCIDENCE, skinning.REPRODUCIBLE = True, False

    # vertex -> face-corner lists from the same builder (many short lists: scatter + per-list sort)
    faces_i = scene.faces.int().to(DEV).contiguous()
    vptr, vinc = skinning.node_incidence(faces_i, scene.verts.shape[0])
    fflat = scene.faces.numpy().reshape(-1)
    assert np.array_equal(vinc.cpu().numpy(), np.argsort(fflat, kind="stable").astype(np.int32))
    assert np.array_equal(vptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(fflat, minlength=scene.verts.shape[0]))]).astype(np.int32))

    a, b, c = grads(True), grads(True), grads(False)        # gather-based twice, list-free once
    for x, y, z in zip(a, b, c):
        # gather-based backward (M * n_t = 1280 >= 1184: unsplit node lists): no floating-point reductions anywhere,
        # so two runs agree bit for bit; the list-free path (reductions) agrees to rounding
        assert torch.equal(x, y)
        assert Hh.rel_linf(x.cpu().double(), z.cpu().double()) <= 1e-4


@pytest.mark.parametrize("n_faces,g,M,K,T", [(1002, 4, 7, 3, 3), (330, 1, 2, 1, 5), (2050, 6, 40, 8, 2)])
def test_skin_ragged_sizes(n_faces, g, M, K, T):
    """Face counts that are no multiple of the warp size (warps straddle two timestamps in the staged Gaussian kernels),
    every Gaussians-per-face pattern, K = 1 and tiny graphs (node lists far longer than a CTA, split over CTAs)."""
    scene = synthetic.make_sugar_scene(n_faces, g=g)
    graph = synthetic.make_deform_graph(scene.verts, M, K)
    node = synthetic.random_node_attrs(T, M, seed=n_faces)
    run_both(scene, graph, node, "hybrid", seed=g)
