"""Multi-GPU plumbing of the path: camera sharding.

One process per GPU, `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests).  The path shards over views with no
data-path collective; the only exchange step is the sum of gradients, which lives with the optimizer step
(`trainstep.FlatGradBucket.all_reduce` / `trainstep.node_attribute_backward`) and with the bench's rasterizer step.
"""
from __future__ import annotations

import torch


def shard_views(n_views: int, rank: int, world: int) -> torch.Tensor:
    """Round-robin view indices of this rank: {rank, rank+world, ...}."""
    return torch.arange(rank, n_views, world)
