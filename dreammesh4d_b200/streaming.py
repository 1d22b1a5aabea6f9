"""Rasterize fwd+bwd for callers whose Gaussian sets and image gradients live in HOST memory.

``HostStreamedRasterStep`` software-pipelines consecutive steps over three CUDA streams with double-buffered
device inputs: the host->device copies of step i+1, the forward+backward of step i (public ``rasterize_batch``
API, all views in one launch sequence) and the device->host copies of step i-1 overlap, so a stream of steps
costs ~max(PCIe time, kernel time) per step instead of their sum.  No host synchronisation inside ``step()``.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, Dict, Optional

import torch

from . import rasterizer as R

SHARED = ("scales", "opac", "cols")


class HostStreamedRasterStep:
    """``host_sets``: pinned tensors ``means [B,P,3]``, ``rots [B,P,4]`` (one attribute set per view) and shared
    ``scales [P,3]``, ``opac [P,1]``, ``cols [P,3]``; ``host_grads``: pinned ``gC [B,3,H,W]``, ``gD``/``gA [B,1,H,W]``.
    ``step()`` enqueues one full step (H2D of every input, forward, backward, optional cross-rank reduction of the
    shared-attribute gradients, D2H of images and gradients into the pinned ``out`` tensors)."""

    def __init__(self, host_sets: Dict[str, torch.Tensor], host_grads: Dict[str, torch.Tensor], view_params: torch.Tensor,
                 H: int, W: int, capacity: int, reduce_fn: Optional[Callable[[torch.Tensor], None]] = None):
        dev = view_params.device
        self.H, self.W, self.capacity, self.reduce_fn = H, W, int(capacity), reduce_fn
        self.hs, self.hg, self.vp = host_sets, host_grads, view_params
        B, P = host_sets["means"].shape[0], host_sets["means"].shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        self.bufs = []
        for _ in range(2):       # double-buffered device inputs
            b = {k: torch.empty(v.shape, **f32).requires_grad_(True) for k, v in host_sets.items()}
            b.update({k: torch.empty(v.shape, **f32) for k, v in host_grads.items()})
            self.bufs.append(b)
        pin = lambda *s: torch.empty(*s).pin_memory()
        self.out = {"means": pin(B, P, 3), "rots": pin(B, P, 4), "scales": pin(P, 3), "opac": pin(P, 1), "cols": pin(P, 3),
                    "color": pin(B, 3, H, W), "depth": pin(B, 1, H, W), "alpha": pin(B, 1, H, W)}
        self.h2d, self.d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.h2d_bytes = sum(v.numel() * 4 for v in host_sets.values()) + sum(v.numel() * 4 for v in host_grads.values())
        self.d2h_bytes = sum(v.numel() * 4 for v in self.out.values())
        self.i = 0
        self.compute_done = [None, None]          # event: compute finished reading buffer k
        self.d2h_done: Optional[torch.cuda.Event] = None
        self.inflight: deque = deque()
        self.retain_all = False                   # set while capturing into a CUDA graph (see run_many)
        self.last_state: Optional[R.RasterState] = None
        # compute runs in stream order (forward i, backward i, forward i+1, ...): one persistent scratch serves all
        # steps, so the eager path allocates no GB-sized blocks per step (the copy streams only touch outputs)
        self.workspace = R.RasterWorkspace()

    def step(self) -> None:
        cur = torch.cuda.current_stream()
        k = self.i & 1
        buf = self.bufs[k]
        # ---- H2D (copy stream): may run while the previous step is still computing on the other buffer ----
        with torch.no_grad(), torch.cuda.stream(self.h2d):
            if self.compute_done[k] is not None:
                self.h2d.wait_event(self.compute_done[k])
            else:
                self.h2d.wait_stream(cur)
            for name, src in self.hs.items():
                buf[name].copy_(src, non_blocking=True)
            for name, src in self.hg.items():
                buf[name].copy_(src, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.h2d)
        # ---- compute (current stream) ----
        cur.wait_event(ev_in)
        st: list = []
        color, radii, depth, alpha = R.rasterize_batch(buf["means"], buf["opac"], buf["scales"], buf["rots"], buf["cols"],
                                                       self.vp, self.H, self.W, capacity=self.capacity, distinct_sets=True,
                                                       state_out=st, workspace=self.workspace)
        torch.autograd.backward([color, depth, alpha], [buf["gC"], buf["gD"], buf["gA"]])
        self.last_state = st[0]
        grads = {n: buf[n].grad for n in self.hs}
        for n in self.hs:
            buf[n].grad = None
        if self.reduce_fn is not None:            # the path's one exchange step (NCCL sum of the shared gradients)
            flat = torch.cat([grads[n].reshape(-1) for n in SHARED])
            self.reduce_fn(flat)
            off = 0
            for n in SHARED:
                cnt = grads[n].numel()
                grads[n] = flat[off:off + cnt].view_as(grads[n])
                off += cnt
        ev_c = torch.cuda.Event()
        ev_c.record(cur)
        self.compute_done[k] = ev_c
        # ---- D2H (copy stream): overlaps the next step's compute ----
        outs = dict(grads, color=color, depth=depth, alpha=alpha)
        with torch.no_grad(), torch.cuda.stream(self.d2h):
            self.d2h.wait_event(ev_c)
            for n, t in outs.items():
                if not self.retain_all:
                    t.record_stream(self.d2h)
                self.out[n].copy_(t, non_blocking=True)
            self.d2h_done = torch.cuda.Event()
            self.d2h_done.record(self.d2h)
        self.inflight.append(outs)
        while not self.retain_all and len(self.inflight) > 2:
            self.inflight.popleft()
        self.i += 1

    def run_many(self, k: int) -> None:
        """``k`` pipelined steps followed by ``drain()``.  Safe to capture into ONE CUDA graph (the copy streams are
        forked from and joined back into the current stream; every output stays referenced until the capture ends,
        so the graph's memory pool never recycles a block that another stream still reads)."""
        capturing = torch.cuda.is_current_stream_capturing()
        self.retain_all = capturing
        self.compute_done = [None, None]
        for _ in range(k):
            self.step()
        self.drain()
        self.retain_all = False
        if not capturing:
            while len(self.inflight) > 2:
                self.inflight.popleft()

    def drain(self) -> None:
        """Makes the current stream wait for every outstanding copy (end of a timed region / before reading ``out``)."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.h2d)
        cur.wait_stream(self.d2h)

    def overflowed(self) -> bool:
        return self.last_state is not None and self.last_state.status()[1]
