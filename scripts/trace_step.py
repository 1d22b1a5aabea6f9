import sys, time, torch, json
sys.path.insert(0, '.')
exec(open('scripts/diag_step.py').read().split("flush = torch.empty")[0])
from torch.profiler import profile, ProfilerActivity
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4):
        flush.zero_()
        step()
    torch.cuda.synchronize()
prof.export_chrome_trace("gpurun_out/trace.json")
ev = json.load(open("gpurun_out/trace.json"))["traceEvents"]
gpu = [e for e in ev if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
gpu.sort(key=lambda e: e["ts"])
t0 = gpu[0]["ts"]
prev_end = t0
for e in gpu:
    gap = e["ts"] - prev_end
    print(f'{e["ts"]-t0:10.1f} us  dur {e["dur"]:9.1f}  gap {gap:8.1f}  {e["cat"]:10s} {e["name"][:60]}')
    prev_end = e["ts"] + e["dur"]
