import sys, time, torch, json, gc
sys.path.insert(0, '.')
exec(open('scripts/diag_step.py').read().split("flush = torch.empty")[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(label, do_flush, n=40):
    for _ in range(3): step()
    torch.cuda.synchronize()
    evs, host = [], []
    for i in range(n):
        if do_flush: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter(); e0.record(); step(); e1.record(); h1 = time.perf_counter()
        evs.append((e0, e1)); host.append((h1 - h0) * 1e3)
    torch.cuda.synchronize()
    g = [a.elapsed_time(b) for a, b in evs]
    print(label, "gpu mean %.2f max %.1f" % (sum(g)/len(g), max(g)), "host mean %.2f max %.1f" % (sum(host)/len(host), max(host)), "n>12ms:", sum(x > 12 for x in g))
print("gc enabled:", gc.isenabled(), gc.get_count(), len(gc.get_objects()))
for r in range(3): run("gc-on  flush", True)
gc.collect(); gc.freeze(); gc.disable()
for r in range(3): run("gc-off flush", True)
import os
print("loadavg", os.getloadavg(), "cpus", os.cpu_count())
