"""HexPlane + MLP deformation network (row A1 of SURVEY.md §8a): M <= 1000 query points per timestamp,
launch-latency bound, feeds the fused skinning kernel.  The plane lookup (6 planes x 4 scales of bilinear samples,
their products and the concatenation) runs as ONE fused kernel each way (``hexplane.py`` -> dm4d_hexplane_*,
row (f)3); the small MLP heads stay PyTorch tensor-core matmuls, as the north-star prescribes for A1.
``fused=False`` evaluates the lookup with F.grid_sample exactly like the reference (used by tests as the PyTorch
statement of A1 and for CPU tensors); on CUDA the default is the fused kernel and nothing falls back silently.

Own implementation of the network the reference builds in
custom/threestudio-dreammesh4d/geometry/deformation.py (HexPlaneField :177-248, interpolate_ms_features
:141-174, Deformation.create_res_net/forward_dynamic_delta :366-436, DeformationNetwork :477-554), with the
same parameter names so the reference's checkpoints load (`_deformation.*`) and the same
quirks (aabb = [[+b],[-b]] => coordinates are negated; zero-initialised residual heads).
All timestamps of a step are evaluated in one batch.
"""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F

PLANES = list(itertools.combinations(range(4), 2))      # (x,y) (x,z) (x,t) (y,z) (y,t) (z,t)


class _HexPlanes(nn.Module):
    def __init__(self, bounds=1.0, feat=32, base_res=(64, 64, 64, 25), multires=(1, 2, 4, 8), fused=True):
        super().__init__()
        self.fused = fused
        self.aabb = nn.Parameter(torch.tensor([[bounds] * 3, [-bounds] * 3]), requires_grad=False)
        self.grids = nn.ModuleList()
        for m in multires:
            res = [r * m for r in base_res[:3]] + [base_res[3]]
            planes = nn.ParameterList()
            for (i, j) in PLANES:
                p = nn.Parameter(torch.empty(1, feat, res[j], res[i]))
                if 3 in (i, j):
                    nn.init.ones_(p)                       # time planes start at 1
                else:
                    nn.init.uniform_(p, a=0.1, b=0.5)
                planes.append(p)
            self.grids.append(planes)
        self.feat_dim = feat * len(multires)

    def forward(self, pts: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """pts [N,3], t [N,1] -> [N, feat_dim]."""
        p = (pts - self.aabb[0]) * (2.0 / (self.aabb[1] - self.aabb[0])) - 1.0
        q = torch.cat([p, t], dim=-1)
        if self.fused:
            from .hexplane import hexplane_features
            return hexplane_features(q, [list(planes) for planes in self.grids])
        feats = []
        for planes in self.grids:
            prod = 1.0
            for plane, (i, j) in zip(planes, PLANES):
                coords = q[:, [i, j]].view(1, 1, -1, 2)
                s = F.grid_sample(plane, coords, mode="bilinear", padding_mode="border", align_corners=True)
                prod = prod * s.view(plane.shape[1], -1).t()
            feats.append(prod)
        return torch.cat(feats, dim=-1)


class _LinearRes(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.main_stream = nn.Linear(w, w)

    def forward(self, x):
        x = F.relu(x)
        return x + self.main_stream(x)


class _Head(nn.Module):
    def __init__(self, w, out):
        super().__init__()
        self.feature_out = nn.Sequential(_LinearRes(w), nn.Linear(w, out))

    def forward(self, x):
        return self.feature_out(x)


class _DeformationNet(nn.Module):
    def __init__(self, width=64, no_ds=False, no_do=False, **grid_kw):
        super().__init__()
        self.grid = _HexPlanes(**grid_kw)
        self.feature_out = nn.Sequential(nn.Linear(self.grid.feat_dim, width))       # defor_depth = 1
        self.pos_deform, self.scales_deform = _Head(width, 3), _Head(width, 6)
        self.rotations_deform, self.opacity_deform = _Head(width, 4), _Head(width, 1)
        self.no_ds, self.no_do = no_ds, no_do


class HexPlaneDeformation(nn.Module):
    """Callable ``(node_xyz [M,3], timestamps [T]) -> (trans [T,M,3], rot_delta [T,M,4], strain [T,M,6] | None,
    opacity_delta [T,M,1] | None)``; timestamps in (0,1) are mapped to 2t-1 (dynamic_sugar.py:431)."""

    def __init__(self, width=64, no_ds=False, no_do=False, timebase_pe=4, timenet_width=64, timenet_output=32, **grid_kw):
        """``grid_kw``: bounds, feat, base_res, multires, fused (see _HexPlanes)."""
        super().__init__()
        self.timenet = nn.Sequential(nn.Linear(2 * timebase_pe + 1, timenet_width), nn.ReLU(),
                                     nn.Linear(timenet_width, timenet_output))     # present but unused, as in the reference
        self.deformation_net = _DeformationNet(width, no_ds, no_do, **grid_kw)
        self.register_buffer("time_poc", torch.tensor([2.0 ** i for i in range(timebase_pe)]))
        self.register_buffer("pos_poc", torch.tensor([2.0 ** i for i in range(10)]))
        self.register_buffer("rotation_scaling_poc", torch.tensor([2.0 ** i for i in range(2)]))
        self.register_buffer("opacity_poc", torch.tensor([2.0 ** i for i in range(2)]))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=1)
        for head in (self.deformation_net.pos_deform, self.deformation_net.scales_deform,
                     self.deformation_net.rotations_deform, self.deformation_net.opacity_deform):
            for m in head.modules():
                if isinstance(m, nn.Linear):
                    nn.init.zeros_(m.weight)
                    nn.init.zeros_(m.bias)

    def forward(self, node_xyz: torch.Tensor, timestamps: torch.Tensor):
        T, M = timestamps.shape[0], node_xyz.shape[0]
        pts = node_xyz.repeat(T, 1)
        t = (timestamps * 2 - 1).repeat_interleave(M)[:, None].to(pts.dtype)
        net = self.deformation_net
        h = net.feature_out(net.grid(pts, t)).float()
        trans = net.pos_deform(h).view(T, M, 3)
        rot = net.rotations_deform(h).view(T, M, 4)
        strain = None if net.no_ds else net.scales_deform(h).view(T, M, 6)
        opac = None if net.no_do else net.opacity_deform(h).view(T, M, 1)
        return trans, rot, strain, opac
