import sys, time, torch, json
sys.path.insert(0, '.')
exec(open('scripts/diag_step.py').read().split("flush = torch.empty")[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(label, do_flush, n=12, sleep=0.0):
    for _ in range(3): step()
    torch.cuda.synchronize()
    evs, host = [], []
    for i in range(n):
        if sleep: time.sleep(sleep)
        if do_flush: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter(); e0.record(); step(); e1.record(); h1 = time.perf_counter()
        evs.append((e0, e1)); host.append((h1 - h0) * 1e3)
    torch.cuda.synchronize()
    print(label, "gpu:", [round(a.elapsed_time(b), 1) for a, b in evs], "host:", [round(h, 1) for h in host])
run("noflush", False)
run("flush", True)
run("noflush", False)
run("flush", True)
run("flush+sleep", True, sleep=0.05)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,pstate --format=csv", shell=True, capture_output=True, text=True).stdout)
