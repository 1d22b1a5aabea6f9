"""Deformation-graph construction — host-side mirror of
DynamicSuGaRModel.build_deformation_graph (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:745-861) for
the Euclidean mode, on the GPU: one brute-force K-nearest-node kernel for all vertices instead of one KD-tree
query, one numpy conversion and one host->device copy per vertex.

Wire format (what the skinning kernels and the reference's checkpoints expect): ``_xyz_neighbor_node_idx [V,K]``
long, ``_xyz_neighbor_nodes_weights [V,K]`` row-normalised, ``_deform_graph_node_xyz [M,3]``, plus the node-node
``_deform_graph_connectivity [M,K]``.

Modes
* ``"eucdisc"``  — the reference's Euclidean mode: weights = the KD-tree's SQUARED distances, row-normalised
  (:783-789, 857-861).
* ``"falloff"``  — Euclidean neighbour sets with the weight formula of the geodesic mode,
  ``(1 - d_k / d_{K+1})^2`` row-normalised (:845, 859-861): the synthetic benchmark graphs.
The geodesic mode itself needs potpourri3d's heat-method solver (un-vendored, sparse Cholesky per mesh) and is not
built.  Control nodes are an input (the reference samples them with Open3D's RNG, or takes ``xyz_nodes``).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .synthetic import DeformGraph


def knn_nodes(queries: torch.Tensor, nodes: torch.Tensor, k: int):
    """(idx [N,k] int32, sqdist [N,k] fp32) of the k nearest ``nodes`` of every query, ascending (distance, index)."""
    if queries.device.type != "cuda":
        raise _lib.Dm4dError("dreammesh4d_b200 graph construction needs CUDA tensors (there is no CPU path)")
    q, n = queries.contiguous().float(), nodes.contiguous().float()
    idx = torch.empty(q.shape[0], k, dtype=torch.int32, device=q.device)
    d2 = torch.empty(q.shape[0], k, dtype=torch.float32, device=q.device)
    check(_lib.lib().dm4d_graph_knn(ptr(q), q.shape[0], ptr(n), n.shape[0], k, ptr(idx), ptr(d2),
                                    torch.cuda.current_stream().cuda_stream), "dm4d_graph_knn")
    return idx, d2


def build_deformation_graph(verts: torch.Tensor, node_xyz: torch.Tensor, nodes_connectivity: int = 6,
                            mode: str = "eucdisc"):
    """Returns ``(DeformGraph, connectivity [M,K] long)`` on the device of ``verts``."""
    K = nodes_connectivity
    if mode == "eucdisc":
        idx, d2 = knn_nodes(verts, node_xyz, K)
        w = d2
    elif mode == "falloff":
        idx, d2 = knn_nodes(verts, node_xyz, K + 1)
        d = d2.sqrt()
        w = (1.0 - d[:, :K] / d[:, K:K + 1].clamp_min(1e-12)) ** 2
        idx = idx[:, :K]
    else:
        raise ValueError("mode must be 'eucdisc' or 'falloff' (the geodesic mode needs potpourri3d and is not built)")
    w = w / w.sum(dim=-1, keepdim=True).clamp_min(1e-12) if mode == "falloff" else w / w.sum(dim=-1, keepdim=True)
    conn, _ = knn_nodes(node_xyz, node_xyz, K + 1)
    return DeformGraph(node_xyz.float(), idx.long(), w.float()), conn[:, 1:].long()
