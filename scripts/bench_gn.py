#!/usr/bin/env python
"""GroupNorm forward / backward timing at the encoder's and the UNet's activation shapes (achieved HBM GB/s)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dreammesh4d_b200.nhwc import groupnorm_nhwc

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for shape in ((8, 128, 256, 256), (8, 256, 128, 128), (8, 512, 64, 64), (8, 512, 32, 32), (16, 320, 32, 32)):
    x = torch.randn(shape, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g, b = torch.ones(shape[1], device=dev), torch.zeros(shape[1], device=dev)
    gy = torch.randn(shape, device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
    res = []
    for mode in ("fwd", "bwd"):
        ts = []
        for i in range(8):
            y = groupnorm_nhwc(x, g, b, 32, 1e-6, True)
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if mode == "fwd":
                e0.record(); y = groupnorm_nhwc(x, g, b, 32, 1e-6, True); e1.record()
            else:
                e0.record(); torch.autograd.grad(y, x, gy); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(sorted(ts)[len(ts) // 2] * 1e3)
    nbytes = x.numel() * 2
    out.append(f"{shape}: fwd {res[0]:.1f} us ({2 * nbytes / res[0] / 1e3:.0f} GB/s of 2 passes), bwd {res[1]:.1f} us ({3 * nbytes / res[1] / 1e3:.0f} GB/s of 3 passes)")
print(" | ".join(out))
