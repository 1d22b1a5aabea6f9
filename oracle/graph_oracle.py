"""CPU oracle of the Euclidean deformation-graph construction.

TEST INFRASTRUCTURE ONLY.  numpy restatement of DynamicSuGaRModel.build_deformation_graph, mode "eucdisc"
(custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:765-790, 853-861): per vertex the K nearest control
nodes and their SQUARED distances (KDTreeFlann.search_knn_vector_3d results [1], [2]), weights = squared distances
normalised per row; node connectivity = K+1 nearest nodes of a node minus the first (itself).
PARITY UNPINNED for the neighbour search itself: Open3D (KDTreeFlann) is not installed here and the reference has no
test for it; an exact K-nearest search is uniquely defined up to ties, which this oracle breaks by node index.
Arithmetic: d2 = (dx*dx + dy*dy) + dz*dz in binary32, every operation rounded (the kernel's spec).
"""
from __future__ import annotations

import numpy as np


def knn(queries: np.ndarray, nodes: np.ndarray, k: int):
    q = queries.astype(np.float32)
    n = nodes.astype(np.float32)
    idx = np.empty((q.shape[0], k), np.int32)
    d2o = np.empty((q.shape[0], k), np.float32)
    for s in range(0, q.shape[0], 4096):
        d = q[s:s + 4096, None, :] - n[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]     # float32, op by op
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]                             # ties -> lower index first
        idx[s:s + 4096] = order
        d2o[s:s + 4096] = np.take_along_axis(d2, order, axis=1)
    return idx, d2o


def build_eucdisc(verts: np.ndarray, nodes: np.ndarray, K: int):
    idx, d2 = knn(verts, nodes, K)
    w = d2 / d2.sum(axis=-1, keepdims=True)
    conn, _ = knn(nodes, nodes, K + 1)
    return idx.astype(np.int64), w.astype(np.float32), conn[:, 1:].astype(np.int64)


def geodesic_knn(verts: np.ndarray, faces: np.ndarray, nodes: np.ndarray, k: int):
    """Mode "geodisc" (custom/threestudio-dreammesh4d/geometry/dynamic_sugar.py:791-849), restated with exact shortest
    paths along the mesh edges: every node enters the mesh at its nearest vertex (:809-815), per vertex the k nodes with
    the smallest path length (ties by node index).  The reference ranks by potpourri3d's heat-method distances
    (un-vendored; PARITY UNPINNED): a smoothed approximation of the same geodesic distance — Dijkstra on the edge graph
    is the exact distance of the piecewise-linear edge metric, which bounds it from above.
    Returns (idx [V,k] int32, dist [V,k] float64)."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import dijkstra
    V = verts.shape[0]
    v64 = verts.astype(np.float32).astype(np.float64)
    i = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2]])
    j = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0]])
    w = np.linalg.norm(verts.astype(np.float32)[i] - verts.astype(np.float32)[j], axis=1).astype(np.float64)   # fp32 edge lengths, as the kernel
    g = coo_matrix((w, (i, j)), shape=(V, V)).tocsr()
    g = g.maximum(g.T)
    entry = knn(nodes, verts, 1)[0][:, 0]
    d = dijkstra(g, directed=False, indices=entry)              # [M, V]
    order = np.argsort(d.T, axis=1, kind="stable")[:, :k]       # ties -> lower node index first
    return order.astype(np.int32), np.take_along_axis(d.T, order, axis=1)


def build_geodisc(verts: np.ndarray, faces: np.ndarray, nodes: np.ndarray, K: int):
    idx, _ = geodesic_knn(verts, faces, nodes, K + 1)
    d = np.linalg.norm(verts[:, None, :].astype(np.float64) - nodes[idx].astype(np.float64), axis=-1)   # Euclidean (:836-838)
    w = (1.0 - d[:, :K] / d[:, K:K + 1]) ** 2
    return idx[:, :K].astype(np.int64), (w / w.sum(-1, keepdims=True)).astype(np.float32)
