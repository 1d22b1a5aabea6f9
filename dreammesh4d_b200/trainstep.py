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
        total = None
        for batch in batches:
            ts = batch["timestamp"]
            with torch.no_grad():
                node = geo.get_timed_dg_attributes(ts)
            node = [None if t is None else t.detach().requires_grad_(True) for t in node]
            out = self.ren.batch_forward(batch, node_attrs=node)
            loss = self.loss_fn(out, batch)
            loss.backward()                                    # ... down to the control-node attributes
            node_attribute_backward(geo._deformation, geo._deform_graph_node_xyz, ts, [None if t is None else t.grad for t in node],
                                    self.group)                # exchange + replicated network backward
            total = loss.detach() if total is None else total + loss.detach()
            geo.update_step(0, step)                           # per-substep caches (dynamic_sugar.py:863-873)
        self.opt.step()
        return total
