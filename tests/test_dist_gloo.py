"""CPU, world_size 2, gloo: host-side logic of the multi-GPU path (view sharding + the single gradient exchange)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dreammesh4d_b200 import dist as D


def _worker(rank, world, port, ret):
    """FlatGradBucket: every .grad is a view of one flat buffer; all_reduce sums it across ranks in one collective."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dreammesh4d_b200.trainstep import FlatGradBucket
        params = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(4)),
                  torch.nn.Parameter(torch.zeros(5), requires_grad=False)]
        bucket = FlatGradBucket(params)
        params[0].grad.fill_(float(rank + 1))
        params[1].grad.fill_(float(10 * (rank + 1)))
        bucket.all_reduce()
        ok = torch.allclose(params[0].grad, torch.full((3, 2), 3.0)) and torch.allclose(params[1].grad, torch.full((4,), 30.0))
        ok = ok and params[0].grad.data_ptr() == bucket.flat.data_ptr() and params[2].grad is None
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_view_sharding_is_a_partition():
    for n, w in ((8, 1), (8, 2), (8, 8), (7, 4)):
        parts = [D.shard_views(n, r, w) for r in range(w)]
        assert sorted(torch.cat(parts).tolist()) == list(range(n))


@pytest.mark.timeout(120)
def test_gradient_exchange_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def _worker_nodegrad(rank, world, port, ret):
    """node_attribute_backward: gathered replay == sum over ranks of the per-rank parameter gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dreammesh4d_b200.deformation import HexPlaneDeformation
        from dreammesh4d_b200.geometry import activate_node_deltas
        from dreammesh4d_b200.trainstep import node_attribute_backward
        torch.manual_seed(0)
        net = HexPlaneDeformation(base_res=(8, 8, 8, 5), multires=(1, 2), fused=False)   # CPU (gloo) test: PyTorch lookup
        with torch.no_grad():
            for head in (net.deformation_net.pos_deform, net.deformation_net.rotations_deform,
                         net.deformation_net.scales_deform, net.deformation_net.opacity_deform):
                head.feature_out[1].weight.normal_(0, 0.1)
        xyz = torch.rand(7, 3) - 0.5
        ts_all = torch.linspace(0.1, 0.9, 2 * world)
        g = torch.Generator().manual_seed(1)
        shapes = [(2 * world, 7, 3), (2 * world, 7, 4), (2 * world, 7, 3, 3), (2 * world, 7, 1)]
        g_all = [torch.randn(s, generator=g) for s in shapes]
        # reference: single-process backward over the global batch
        attrs = activate_node_deltas(*net(xyz, ts_all))
        torch.autograd.backward(list(attrs), g_all)
        want = [p.grad.clone() for p in net.parameters() if p.grad is not None]
        net.zero_grad(set_to_none=True)
        # distributed: every rank holds its slice of timestamps / node gradients
        sl = slice(2 * rank, 2 * rank + 2)
        node_attribute_backward(net, xyz, ts_all[sl], [t[sl] for t in g_all])
        got = [p.grad for p in net.parameters() if p.grad is not None]
        ret[rank] = len(got) == len(want) and all(torch.allclose(a, b, rtol=1e-5, atol=1e-7) for a, b in zip(got, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_node_attribute_gradient_exchange_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_nodegrad, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


class _StubGeometry:
    """The slice of the geometry interface DynamicStageStep touches, around a real (PyTorch-lookup) deformation net."""

    def __init__(self, net, xyz):
        self._deformation, self._deform_graph_node_xyz = net, xyz

    def update_step(self, *a, **k):
        pass

    def get_timed_dg_attributes(self, ts):
        from dreammesh4d_b200.geometry import activate_node_deltas
        return activate_node_deltas(*self._deformation(self._deform_graph_node_xyz, ts))


class _StubRenderer:
    """A differentiable stand-in for skinning + rasterizer + post-ops: any smooth function of the node attributes."""

    def batch_forward(self, batch, node_attrs=None):
        trans, rot, scale, opac = node_attrs
        return {"x": (trans * batch["w"][:, None, None]).sum() + (rot ** 2 * batch["w"][:, None, None]).sum() +
                     (scale.sum(dim=(-1, -2)) * batch["w"][:, None]).sum() * 0.1 + (opac[..., 0] * batch["w"][:, None]).sum()}


def _make_net():
    from dreammesh4d_b200.deformation import HexPlaneDeformation
    torch.manual_seed(0)
    net = HexPlaneDeformation(base_res=(8, 8, 8, 5), multires=(1, 2), fused=False)
    with torch.no_grad():
        for head in (net.deformation_net.pos_deform, net.deformation_net.rotations_deform,
                     net.deformation_net.scales_deform, net.deformation_net.opacity_deform):
            head.feature_out[1].weight.normal_(0, 0.1)
    return net


def _worker_step(rank, world, port, ret, exchange):
    """DynamicStageStep with a process group: every rank steps on ITS batch; the parameters end up identical on all
    ranks and equal to a single-process step over the union of the batches (the reference's DDP semantics)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dreammesh4d_b200.trainstep import DynamicStageStep
    xyz = torch.rand(7, 3, generator=torch.Generator().manual_seed(5)) - 0.5
    g = torch.Generator().manual_seed(9)
    batches = [{"timestamp": torch.rand(3, generator=g), "w": torch.randn(3, generator=g)} for _ in range(world)]
    loss_fn = lambda out, b: out["x"]
    # single-process reference over the union (no process group yet)
    net_ref = _make_net()
    ref_step = DynamicStageStep(_StubGeometry(net_ref, xyz), _StubRenderer(), torch.optim.SGD(net_ref.parameters(), lr=0.1), loss_fn)
    ref_step(batches, 0)
    ref_step(batches, 1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _make_net()
        step = DynamicStageStep(_StubGeometry(net, xyz), _StubRenderer(), torch.optim.SGD(net.parameters(), lr=0.1), loss_fn,
                                exchange=exchange)
        assert (step.bucket is not None) == (exchange == "dense")
        step([batches[rank]], 0)
        step([batches[rank]], 1)          # second step: the flat bucket is re-zeroed, not re-accumulated
        moved = sum(float((p - q).abs().max()) > 0 for p, q in zip(net.parameters(), _make_net().parameters()))
        same = all(torch.allclose(p, q, rtol=1e-5, atol=1e-7) for p, q in zip(net.parameters(), net_ref.parameters()))
        ret[rank] = bool(same and moved >= 8)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("exchange", ["dense", "node_gather"])
def test_dynamic_stage_step_world2_gloo_equals_single_process_union(exchange):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() % 2000) + (7 if exchange == "dense" else 0)
    mp.spawn(_worker_step, args=(2, port, ret, exchange), nprocs=2, join=True)
    assert ret[0] and ret[1]
