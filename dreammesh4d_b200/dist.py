"""Multi-GPU plumbing of the path: camera sharding + the one gradient exchange.

One process per GPU, `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests).  The path shards over
views with no data-path collective; the only exchange step is the sum of gradients — preferably of the
control-node attributes (n_frames x M x 17 floats), which the skinning backward produces directly.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> torch.Tensor:
    """Round-robin view indices of this rank: {rank, rank+world, ...}."""
    return torch.arange(rank, n_views, world)


def views_of_all_ranks(n_views: int, world: int) -> List[torch.Tensor]:
    return [shard_views(n_views, r, world) for r in range(world)]


def allreduce_sum_(tensors: Iterable[Optional[torch.Tensor]], group=None) -> None:
    """In-place SUM all-reduce of a list of tensors as ONE flat bucket (one collective per step)."""
    ts = [t for t in tensors if t is not None]
    if not ts or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in ts:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def node_gradient_exchange(node_grads: Iterable[torch.Tensor], frame_ids: torch.Tensor, n_frames_total: int, group=None):
    """Sums control-node attribute gradients over ranks when every rank rendered a different subset of frames.
    node_grads: per-attribute tensors [T_local, M, k]; frame_ids [T_local] (global frame index of each local
    timestamp).  Returns the per-attribute gradients of ALL frames [n_frames_total, M, k], identical on every rank,
    so each rank can run the (replicated) deformation-network backward locally."""
    outs = []
    for g in node_grads:
        full = torch.zeros(n_frames_total, *g.shape[1:], dtype=g.dtype, device=g.device)
        full.index_add_(0, frame_ids.to(g.device), g)
        outs.append(full)
    allreduce_sum_(outs, group)
    return outs
