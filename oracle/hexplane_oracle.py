"""CPU oracle of the HexPlane multi-scale feature lookup.

TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else; the product path
(dreammesh4d_b200/hexplane.py -> libdm4d.so) never touches it.

Restates, without calling F.grid_sample, what
  custom/threestudio-dreammesh4d/geometry/deformation.py:80-81  (normalize_aabb),
  :84-111 (grid_sample_wrapper: bilinear, padding_mode='border', align_corners=True),
  :141-174 (interpolate_ms_features: product over the 6 coordinate planes, concatenation over scales) and
  :224-248 (HexPlaneField.get_density / forward)
compute.  Plane (i, j) of itertools.combinations(range(4), 2) is stored [1, F, res[j], res[i]]: coordinate i indexes
the last (width) axis, coordinate j the height axis.  Pinned: tests/golden/hexplane.npz holds inputs, outputs and
plane gradients produced by executing the reference's own HexPlaneField (tests/golden/make_hexplane_golden.py);
tests/test_hexplane.py checks this file against them.  Gradients come from autograd (index_put accumulate).
"""
from __future__ import annotations

import itertools
from typing import Sequence

import torch

PLANES = list(itertools.combinations(range(4), 2))


def normalize_aabb(pts: torch.Tensor, aabb: torch.Tensor) -> torch.Tensor:
    return (pts - aabb[0]) * (2.0 / (aabb[1] - aabb[0])) - 1.0


def bilinear_border(plane: torch.Tensor, u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """plane [F,Hh,Ww]; u (width coordinate), v (height coordinate) in normalised [-1,1] units, any range.
    align_corners=True: pixel = (c + 1) / 2 * (size - 1); 'border': the pixel coordinate is clamped to [0, size-1]."""
    Fd, Hh, Ww = plane.shape
    x = ((u + 1.0) * 0.5 * (Ww - 1)).clamp(0, Ww - 1)
    y = ((v + 1.0) * 0.5 * (Hh - 1)).clamp(0, Hh - 1)
    x0, y0 = torch.floor(x), torch.floor(y)
    wx1, wy1 = x - x0, y - y0
    wx0, wy0 = 1.0 - wx1, 1.0 - wy1
    x0i, y0i = x0.long(), y0.long()
    x1i, y1i = (x0i + 1).clamp_max(Ww - 1), (y0i + 1).clamp_max(Hh - 1)     # weight is 0 whenever the clamp acts
    g = lambda yi, xi: plane[:, yi, xi]                                      # [F,N]
    out = g(y0i, x0i) * (wy0 * wx0) + g(y0i, x1i) * (wy0 * wx1) + g(y1i, x0i) * (wy1 * wx0) + g(y1i, x1i) * (wy1 * wx1)
    return out.t()                                                           # [N,F]


def hexplane_features(coords4: torch.Tensor, grids: Sequence[Sequence[torch.Tensor]]) -> torch.Tensor:
    """coords4 [N,4] normalised (x,y,z,t); grids[s][p] = plane p of scale s, [1,F,res_j,res_i].  Returns [N, S*F]."""
    feats = []
    for planes in grids:
        prod = 1.0
        for plane, (i, j) in zip(planes, PLANES):
            prod = prod * bilinear_border(plane[0], coords4[:, i], coords4[:, j])
        feats.append(prod)
    return torch.cat(feats, dim=-1)


def hexplane_field(pts: torch.Tensor, timestamps: torch.Tensor, aabb: torch.Tensor, grids) -> torch.Tensor:
    """HexPlaneField.forward: pts [N,3] world, timestamps [N,1] already in [-1,1]."""
    return hexplane_features(torch.cat([normalize_aabb(pts, aabb), timestamps], dim=-1), grids)
