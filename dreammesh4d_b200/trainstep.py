"""Dynamic-stage training step on the hot path (mirror of SuGaR4DGen.training_step without the diffusion
guidance, custom/threestudio-dreammesh4d/system/sugar_4dgen.py:397-429): deformation network -> fused skinning
-> batched 6-channel rasterizer -> post-ops -> image losses -> backward -> optimizer.

Multi-GPU: every rank renders its own views at its own timestamps.  Instead of all-reducing the
143 MB of HexPlane gradients, the ranks exchange the gradients of the control-node attributes
(T_local x M x 17 floats, <= 544 KB at M = 1000) together with their timestamps, and every rank replays the (tiny)
deformation-network forward+backward for ALL timestamps locally — identical parameter gradients everywhere, no
parameter-sized collective.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Sequence

import torch
import torch.distributed as dist

from .camera import get_cam_info_gaussian
from .geometry import DynamicSuGaRGeometry, activate_node_deltas
from .renderer import DiffGaussianBatchRenderer


def _gather_cat(t: torch.Tensor, group=None) -> torch.Tensor:
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(outs, t.contiguous(), group=group)
    return torch.cat(outs, dim=0)


def node_attribute_backward(deformation: Callable, node_xyz: torch.Tensor, timestamps: torch.Tensor,
                            node_grads: Sequence[Optional[torch.Tensor]], group=None) -> None:
    """Accumulates d loss / d (deformation-network parameters) for the GLOBAL batch into ``.grad``:
    all-gathers (timestamps, node-attribute gradients) over the ranks and back-propagates them through a local
    replay of the network on all timestamps.  ``node_grads`` = gradients w.r.t. (trans, rot, scale, opacity) as
    returned by ``activate_node_deltas``; entries may be None."""
    ts_all = _gather_cat(timestamps, group)
    attrs = activate_node_deltas(*deformation(node_xyz, ts_all))
    outs, grads = [], []
    for a, g in zip(attrs, node_grads):
        if a is None or g is None:
            continue
        outs.append(a)
        grads.append(_gather_cat(g, group).reshape(a.shape))
    torch.autograd.backward(outs, grads)


class DynamicStageStep:
    """One optimizer step of the dynamic stage on ``batches`` (e.g. the random-camera batch and the reference-camera
    batch of sugar_4dgen.py:411-417).  ``loss_fn(out, batch) -> scalar`` consumes the renderer's output dict."""

    def __init__(self, geometry: DynamicSuGaRGeometry, renderer: DiffGaussianBatchRenderer,
                 optimizer: torch.optim.Optimizer, loss_fn: Callable[[Dict, Dict], torch.Tensor], group=None):
        self.geo, self.ren, self.opt, self.loss_fn, self.group = geometry, renderer, optimizer, loss_fn, group

    def __call__(self, batches: Sequence[Dict], step: int = 0) -> torch.Tensor:
        geo = self.geo
        geo.update_step(0, step)
        self.opt.zero_grad(set_to_none=True)
        # The deformation network is evaluated ONCE per optimizer step for the timestamps of all substeps and
        # back-propagated ONCE with the node-attribute gradients of all substeps: one HexPlane lookup / backward
        # (one zero-fill of the 143 MB of plane gradients, no gradient accumulation passes) instead of one per substep.
        ts_all = torch.cat([b["timestamp"] for b in batches])
        single = not (dist.is_initialized() and dist.get_world_size(self.group) > 1)
        if single:
            live = geo.get_timed_dg_attributes(ts_all)             # autograd graph kept: no replay needed
        else:
            with torch.no_grad():
                live = geo.get_timed_dg_attributes(ts_all)
        leaves = [None if t is None else t.detach().requires_grad_(True) for t in live]
        total, off = None, 0
        for batch in batches:
            n = batch["timestamp"].shape[0]
            node = [None if t is None else t[off:off + n] for t in leaves]
            out = self.ren.batch_forward(batch, node_attrs=node)
            loss = self.loss_fn(out, batch)
            loss.backward()                                    # ... down to the control-node attributes
            total = loss.detach() if total is None else total + loss.detach()
            off += n
            geo.update_step(0, step)                           # per-substep caches (dynamic_sugar.py:863-873)
        grads = [None if t is None else t.grad for t in leaves]
        if single:
            pairs = [(a, g) for a, g in zip(live, grads) if a is not None and g is not None]
            torch.autograd.backward([a for a, _ in pairs], [g for _, g in pairs])
        else:                                                  # exchange + replicated network backward
            node_attribute_backward(geo._deformation, geo._deform_graph_node_xyz, ts_all, grads, self.group)
        self.opt.step()
        return total


class GraphedDynamicStageStep:
    """``DynamicStageStep`` captured into ONE CUDA graph: deformation network, fused skinning, rasterizer, post-ops,
    losses, the whole backward and the optimizer update of every substep replay as a single launch, so the ~600
    kernel launches of a step cost no host time (eager: the step is host/launch-bound, ~3x slower than its kernels).

    Requirements: ``renderer.capacity`` is an integer (no ``num_rendered`` read-back), the optimizer was built with
    ``capturable=True``, ``loss_fn`` has no host synchronisation, and every step uses batches of the shapes of
    ``example_batches``.  Per step the tensor entries of the batches (cameras, timestamps, rays, targets) are copied
    into the graph's static inputs; the camera matrices are derived eagerly before the replay (a batched 4x4
    inverse is a library call that must not be captured).  Single-process: with a process group the node-gradient
    exchange stays outside graphs, use ``DynamicStageStep``."""

    def __init__(self, step: DynamicStageStep, example_batches: Sequence[Dict], warmup: int = 3):
        if step.ren.capacity is None:
            raise ValueError("GraphedDynamicStageStep needs renderer.capacity (an integer) to be set")
        if dist.is_initialized() and dist.get_world_size(step.group) > 1:
            raise ValueError("GraphedDynamicStageStep is single-process; use DynamicStageStep with a process group")
        self.step = step
        self.static = [self._with_cam({k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()})
                       for b in example_batches]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up steps are real optimizer steps
            for i in range(warmup):
                step(self.static, i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step(self.static, warmup)

    @staticmethod
    def _with_cam(b: Dict) -> Dict:
        b["cam_info"] = tuple(t.contiguous() for t in get_cam_info_gaussian(b["c2w"], b["fovy"], b["fovy"], znear=0.1, zfar=100.0))
        return b

    def __call__(self, batches: Sequence[Dict]) -> torch.Tensor:
        for st, b in zip(self.static, batches):
            for k, v in b.items():
                if torch.is_tensor(v):
                    st[k].copy_(v, non_blocking=True)
            cam = get_cam_info_gaussian(st["c2w"], st["fovy"], st["fovy"], znear=0.1, zfar=100.0)
            for dst, src in zip(st["cam_info"], cam):
                dst.copy_(src)
        self.graph.replay()
        return self.loss
