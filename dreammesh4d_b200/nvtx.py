"""NVTX ranges around the phases of the hot path (the reference has no tracing at all — only commented
``time.time_ns()`` prints, dynamic_sugar.py:430-440).  A range costs ~1 us when no profiler is attached; under ``nsys`` /
``ncu --nvtx`` the step reads as  step > substep[k] > {deformation, skin, rasterize, postops, loss, backward} > exchange >
optimizer."""
from __future__ import annotations

import contextlib

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
