"""Dynamic-stage training step on the hot path (mirror of SuGaR4DGen.training_step,
custom/threestudio-dreammesh4d/system/sugar_4dgen.py:397-429, substeps :125-329): deformation network -> fused
skinning -> batched 6-channel rasterizer -> post-ops -> losses (the Zero123 SDS term through ``sds.py`` on the
random-camera substep, image / mesh terms on the reference-camera substep) -> backward -> gradient exchange ->
optimizer.

Multi-GPU: every rank renders its own views at its own timestamps; ONE exchange per step.
Two interchangeable forms, both leave identical parameter gradients on every rank:

  * ``exchange="dense"``   — each rank back-propagates its own node-attribute gradients through the deformation
    network into ONE flat fp32 gradient bucket (the parameters' ``.grad`` are views of it, like a DDP bucket) and the
    bucket is summed with a single NCCL all-reduce (143 MB over NVLink; constant work per rank as N grows).
  * ``exchange="node_gather"`` — ranks all-gather (timestamps, node-attribute gradients: T_local x M x 17 floats,
    <= 544 KB at M = 1000) and every rank replays the deformation-network forward + backward for ALL timestamps
    (no parameter-sized collective, but the replicated network work grows with N).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from .camera import get_cam_info_gaussian
from .geometry import DynamicSuGaRGeometry, activate_node_deltas
from .nvtx import nvtx_range
from .renderer import DiffGaussianBatchRenderer


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def _gather_cat(t: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather along dim 0 into ONE preallocated tensor (a single collective, no Python-side list of tensors)."""
    w = _world(group)
    if w == 1:
        return t
    t = t.contiguous()
    out = torch.empty((w * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out


def node_attribute_backward(deformation: Callable, node_xyz: torch.Tensor, timestamps: torch.Tensor,
                            node_grads: Sequence[Optional[torch.Tensor]], group=None) -> None:
    """``exchange="node_gather"``: accumulates d loss / d (deformation-network parameters) for the GLOBAL batch into
    ``.grad``: all-gathers (timestamps, node-attribute gradients) over the ranks and back-propagates them through a
    local replay of the network on all timestamps.  ``node_grads`` = gradients w.r.t. (trans, rot, scale, opacity) as
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


class FlatGradBucket:
    """One flat fp32 buffer holding the gradients of ``params``; every ``p.grad`` is a view of it, so zeroing is one
    memset and the cross-rank sum is ONE all-reduce with no pack/unpack copies."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        self.attach()

    def attach(self) -> None:
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()
        base = self.flat.untyped_storage().data_ptr()
        if any(p.grad is None or p.grad.untyped_storage().data_ptr() != base for p in self.params):
            self.attach()          # someone called zero_grad(set_to_none=True) or replaced a .grad

    def all_reduce(self, group=None) -> None:
        if _world(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


class DynamicStageStep:
    """One optimizer step of the dynamic stage on ``batches`` (e.g. the random-camera batch and the reference-camera
    batch of sugar_4dgen.py:411-417).  ``loss_fn(out, batch) -> scalar`` consumes the renderer's output dict (the SDS
    term is ``guidance(out["comp_rgb"], **batch)["loss_sds"]``, see ``sds.TemporalStableZero123SDS``).

    Fixed-capacity binning (``renderer.capacity`` set, as CUDA-graph replay requires) cannot grow its workspace on the
    device: if a step's instance count exceeds the capacity the rasterizer renders background and returns zero
    gradients.  The step therefore accumulates the rasterizer's device-side overflow flag (no sync) and
    ``check_overflow()`` raises; in eager mode it is polled every ``overflow_check_every`` steps."""

    def __init__(self, geometry: DynamicSuGaRGeometry, renderer: DiffGaussianBatchRenderer,
                 optimizer: torch.optim.Optimizer, loss_fn: Callable[[Dict, Dict], torch.Tensor], group=None,
                 exchange: str = "dense", overflow_check_every: int = 50):
        if exchange not in ("dense", "node_gather"):
            raise ValueError("exchange must be 'dense' or 'node_gather'")
        self.geo, self.ren, self.opt, self.loss_fn, self.group = geometry, renderer, optimizer, loss_fn, group
        self.exchange = exchange
        self.overflow_check_every = int(overflow_check_every)
        self.overflow_seen: Optional[torch.Tensor] = None
        self._calls = 0
        self.bucket: Optional[FlatGradBucket] = None
        if exchange == "dense" and _world(group) > 1:
            self.bucket = FlatGradBucket([p for grp in optimizer.param_groups for p in grp["params"]])

    # ---- overflow surfacing ------------------------------------------------------------------------------------
    def _note_overflow(self) -> None:
        st = getattr(self.ren, "last_state", None)
        if st is None or getattr(self.ren, "capacity", None) is None:
            return                                   # exact sizing (one read-back per batch) cannot overflow
        flag = st.overflow_flag()
        if self.overflow_seen is None:
            self.overflow_seen = torch.zeros(1, dtype=torch.int32, device=flag.device)
        self.overflow_seen.add_(flag)

    def check_overflow(self) -> None:
        """Synchronises; raises if any step since the last check exceeded ``renderer.capacity``."""
        if self.overflow_seen is not None and int(self.overflow_seen.item()) > 0:
            self.overflow_seen.zero_()
            raise _lib.Dm4dError(
                f"rasterizer bin capacity ({self.ren.capacity} instances) was exceeded: the affected steps rendered "
                "background and produced zero gradients. Raise renderer.capacity (or set it to None for exact sizing) "
                "and, for a captured step, re-capture.")

    # ---- the step ----------------------------------------------------------------------------------------------
    def __call__(self, batches: Sequence[Dict], step: int = 0) -> torch.Tensor:
        geo = self.geo
        geo.update_step(0, step)
        if self.bucket is not None:
            self.bucket.zero()
        else:
            self.opt.zero_grad(set_to_none=True)
        # The deformation network is evaluated ONCE per optimizer step for the timestamps of all substeps and
        # back-propagated ONCE with the node-attribute gradients of all substeps: one HexPlane lookup / backward
        # (one zero-fill of the 143 MB of plane gradients, no gradient accumulation passes) instead of one per substep.
        ts_all = torch.cat([b["timestamp"] for b in batches])
        multi = _world(self.group) > 1
        replay = multi and self.exchange == "node_gather"
        if replay:
            with torch.no_grad():
                live = geo.get_timed_dg_attributes(ts_all)
        else:
            live = geo.get_timed_dg_attributes(ts_all)             # autograd graph kept: no replay needed
        leaves = [None if t is None else t.detach().requires_grad_(True) for t in live]
        total, off = None, 0
        for bi, batch in enumerate(batches):
            n = batch["timestamp"].shape[0]
            node = [None if t is None else t[off:off + n] for t in leaves]
            with nvtx_range(f"dm4d.substep{bi}.forward"):
                out = self.ren.batch_forward(batch, node_attrs=node)
            self._note_overflow()
            with nvtx_range(f"dm4d.substep{bi}.loss"):
                loss = self.loss_fn(out, batch)
            with nvtx_range(f"dm4d.substep{bi}.backward"):
                loss.backward()                                # ... down to the control-node attributes
            total = loss.detach() if total is None else total + loss.detach()
            off += n
            geo.update_step(0, step)                           # per-substep caches (dynamic_sugar.py:863-873)
        grads = [None if t is None else t.grad for t in leaves]
        if replay:                                             # exchange + replicated network backward
            node_attribute_backward(geo._deformation, geo._deform_graph_node_xyz, ts_all, grads, self.group)
        else:
            pairs = [(a, g) for a, g in zip(live, grads) if a is not None and g is not None]
            torch.autograd.backward([a for a, _ in pairs], [g for _, g in pairs])
            if multi:
                with nvtx_range("dm4d.exchange"):
                    self.bucket.all_reduce(self.group)         # the step's one exchange
        with nvtx_range("dm4d.optimizer"):
            self.opt.step()
        self._calls += 1
        if self.overflow_check_every > 0 and self._calls % self.overflow_check_every == 0 and \
                not torch.cuda.is_current_stream_capturing():
            self.check_overflow()
        return total


class GraphedDynamicStageStep:
    """``DynamicStageStep`` captured into ONE CUDA graph: deformation network, fused skinning, rasterizer, post-ops,
    losses (incl. the Zero123 UNet / encoder of the SDS term), the whole backward, the NCCL exchange and the optimizer
    update replay as a single launch, so the ~2000 kernel launches of a step cost no host time.

    Requirements: ``renderer.capacity`` is an integer (no ``num_rendered`` read-back), the optimizer was built with
    ``capturable=True``, ``loss_fn`` has no host synchronisation, and every step uses batches of the shapes of
    ``example_batches``.  Per step the tensor entries of the batches (cameras, timestamps, rays, targets) are copied
    into the graph's static inputs; the camera matrices are derived eagerly before the replay (a batched 4x4
    inverse is a library call that must not be captured).  A replayed graph cannot grow the binning workspace: the
    overflow flag is accumulated on the device and polled every ``overflow_check_every`` replays (raises)."""

    def __init__(self, step: DynamicStageStep, example_batches: Sequence[Dict], warmup: int = 3,
                 overflow_check_every: int = 50):
        if step.ren.capacity is None:
            raise ValueError("GraphedDynamicStageStep needs renderer.capacity (an integer) to be set")
        self.step = step
        self.overflow_check_every = int(overflow_check_every)
        self._replays = 0
        step.overflow_check_every = 0                      # polled here, outside the capture
        self.static = [self._with_cam({k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.items()})
                       for b in example_batches]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up steps are real optimizer steps
            for i in range(warmup):
                step(self.static, i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        step.check_overflow()
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
        self._replays += 1
        if self.overflow_check_every > 0 and self._replays % self.overflow_check_every == 0:
            self.step.check_overflow()
        return self.loss
