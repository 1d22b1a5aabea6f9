#!/usr/bin/env python
"""Kernel-level profile of the two Zero123 pieces of the SDS step (torch.profiler, CUDA time by kernel)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from dreammesh4d_b200 import zero123
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
m = zero123.build_random(device=dev, dtype=torch.float16)
views = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 2 * views
x = torch.randn(n, 4, 32, 32, device=dev, dtype=torch.float16)
t = torch.randint(20, 500, (n,), device=dev)
cond = {"c_concat": [torch.randn(n, 4, 32, 32, device=dev, dtype=torch.float16)], "c_crossattn": [torch.randn(n, 1, 768, device=dev, dtype=torch.float16)]}
img = torch.rand(views, 3, 256, 256, device=dev, requires_grad=True)
g_lat = torch.randn(views, 4, 32, 32, device=dev)


def unet():
    with torch.no_grad():
        return m.apply_model(x, t, cond)


def enc():
    lat = m.get_first_stage_encoding(m.encode_first_stage((img * 2 - 1).half())).float()
    return torch.autograd.grad(lat, img, g_lat)


for name, fn in (("unet_fwd", unet), ("encoder_fwd_bwd", enc)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
    print("=" * 30, name, "(3 iterations)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=90))
