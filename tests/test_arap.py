"""ARAP regulariser: oracle vs the golden vector made by executing the reference's ARAPCoach (CPU), product-side
setup code vs the oracle (CPU), and the fused kernel vs the oracle (GPU)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from dreammesh4d_b200 import arap as A
from oracle import arap_oracle as AO
from tests import helpers as Hh

GOLD = Path(__file__).resolve().parent / "golden" / "arap.npz"


def load():
    z = np.load(GOLD)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_arap_oracle_matches_reference_code():
    z = load()
    e = AO.arap_energy(z["verts"].double(), z["faces"], z["verts_def"].double(), z["rot_xyzw"].double())
    assert Hh.rel_linf(e.numpy(), z["energy"].double().numpy()) <= 1e-5      # the fixture is fp32
    # per-vertex weight sums agree with the reference's [V, max_n] table
    W = AO.cot_weight_matrix(z["verts"].double(), z["faces"])
    nb = AO.one_ring(z["faces"], z["verts"].shape[0])
    mine = torch.stack([W[i, nb[i]].sum() for i in range(len(nb))])
    assert Hh.rel_linf(mine.numpy(), z["edge_cot_weights"].double().sum(dim=1).numpy()) <= 1e-5


def test_cotangent_edge_weights_csr_matches_oracle_matrix():
    z = load()
    row_ptr, col, w = A.cotangent_edge_weights(z["verts"], z["faces"])
    W = AO.cot_weight_matrix(z["verts"].double(), z["faces"])
    V = z["verts"].shape[0]
    dense = torch.zeros(V, V, dtype=torch.float64)
    rows = torch.repeat_interleave(torch.arange(V), (row_ptr[1:] - row_ptr[:-1]).long())
    dense[rows, col.long()] = w.double()
    assert (dense - W).abs().max() <= 1e-5 * W.abs().max()
    nb = AO.one_ring(z["faces"], V)
    assert [sorted(col[row_ptr[i]:row_ptr[i + 1]].tolist()) for i in range(V)] == nb


@pytest.mark.gpu
def test_arap_kernel_matches_oracle_energy_and_gradients():
    z = load()
    dev = "cuda"
    rest, faces = z["verts"], z["faces"]
    xp = z["verts_def"].double().requires_grad_(True)
    q = z["rot_xyzw"].double().requires_grad_(True)
    e_ref = AO.arap_energy(rest.double(), faces, xp, q)
    gE = torch.tensor([0.7, -1.3], dtype=torch.float64)
    (e_ref * gE).sum().backward()
    energy_fn = A.ARAPEnergy(rest.to(dev), faces.to(dev))
    xv = z["verts_def"].to(dev).requires_grad_(True)
    qv = z["rot_xyzw"].to(dev).requires_grad_(True)
    e = energy_fn(xv, qv)
    assert Hh.rel_linf(e.detach().cpu().double().numpy(), e_ref.detach().numpy()) <= 1e-4
    assert Hh.rel_linf(e.detach().cpu().double().numpy(), z["energy"].double().numpy()) <= 1e-4     # reference's own number
    (e * gE.float().to(dev)).sum().backward()
    assert Hh.rel_linf(xv.grad.cpu().double().numpy(), xp.grad.numpy()) <= Hh.TOL_GRAD
    assert Hh.rel_linf(qv.grad.cpu().double().numpy(), q.grad.numpy()) <= Hh.TOL_GRAD


def test_face_pairs_cover_every_interior_edge_once():
    from dreammesh4d_b200 import synthetic
    verts, faces = synthetic.uv_sphere(264)
    pairs = A.face_pairs(faces)
    assert pairs.shape == (264 * 3 // 2, 4)                 # closed manifold: E = 3F/2 edges, one pair each
    assert (pairs[:, 0] < pairs[:, 1]).all()


@pytest.mark.gpu
def test_mesh_normal_consistency_kernel_matches_oracle():
    z = load()
    dev = "cuda"
    xp = z["verts_def"].double().requires_grad_(True)
    ref = AO.mesh_normal_consistency(xp, z["faces"])
    ref.backward()
    pairs = A.face_pairs(z["faces"]).to(dev)
    xv = z["verts_def"].to(dev).requires_grad_(True)
    got = A.mesh_normal_consistency(xv, pairs)
    assert abs(float(got) - float(ref)) <= 1e-5 * max(abs(float(ref)), 1e-6) + 1e-7
    got.backward()
    assert Hh.rel_linf(xv.grad.cpu().double().numpy(), xp.grad.numpy()) <= Hh.TOL_GRAD
