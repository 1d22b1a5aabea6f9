"""CPU, world_size 2, gloo: host-side logic of the multi-GPU path (view sharding + the single gradient exchange)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dreammesh4d_b200 import dist as D


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_frames, M = 6, 5
        mine = D.shard_views(n_frames, rank, world)
        g = torch.Generator().manual_seed(100)
        all_t = torch.randn(n_frames, M, 3, generator=g)
        all_r = torch.randn(n_frames, M, 4, generator=g)
        out = D.node_gradient_exchange([all_t[mine] * (rank + 1), all_r[mine] * (rank + 1)], mine, n_frames)
        scale = torch.tensor([float(1 + (f % world)) for f in range(n_frames)])[:, None, None]
        ok = torch.allclose(out[0], all_t * scale) and torch.allclose(out[1], all_r * scale)
        a, b = torch.full((3, 2), float(rank + 1)), torch.full((4,), float(10 * (rank + 1)))
        D.allreduce_sum_([a, None, b])
        ok = ok and torch.allclose(a, torch.full((3, 2), 3.0)) and torch.allclose(b, torch.full((4,), 30.0))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_view_sharding_is_a_partition():
    for n, w in ((8, 1), (8, 2), (8, 8), (7, 4)):
        parts = D.views_of_all_ranks(n, w)
        assert sorted(torch.cat(parts).tolist()) == list(range(n))


@pytest.mark.timeout(120)
def test_gradient_exchange_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
