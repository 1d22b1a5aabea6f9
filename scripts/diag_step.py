import sys, time, torch
sys.path.insert(0, '.')
import bench
from dreammesh4d_b200 import _lib, rasterizer as R
dev = torch.device('cuda', 0)
scene, graph, node = bench.build_scene(False)
cams = bench.build_cameras(0)
gs = bench.gaussian_sets_gpu(scene, graph, node, dev)
V, PV, campos, tanx, tany = cams
H = W = 512
vp = R.make_view_params(V.to(dev), PV.to(dev), campos.to(dev), tanx, tany, torch.ones(8, 3, device=dev), set_index=torch.arange(8))
inp = {k: gs[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "colors")}
gC = torch.randn(8, 3, H, W, device=dev); gD = torch.randn(8, 1, H, W, device=dev); gA = torch.randn(8, 1, H, W, device=dev)
st = []
with torch.no_grad():
    R.rasterize_batch(inp["means3D"], inp["opacities"], inp["scales"], inp["rotations"], inp["colors"], vp, H, W, state_out=st)
n, _ = st[0].status(); cap = int(n * 1.25) + 4096
def step():
    c, r, d, a = R.rasterize_batch(inp["means3D"], inp["opacities"], inp["scales"], inp["rotations"], inp["colors"], vp, H, W, capacity=cap, distinct_sets=True)
    torch.autograd.backward([c, d, a], [gC, gD, gA])
    for k in inp: inp[k].grad = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(label, prof, do_flush, n=10):
    _lib.profile_enable(prof); _lib.profile_collect()
    for _ in range(3): step()
    torch.cuda.synchronize()
    evs = []
    t0 = time.perf_counter()
    for i in range(n):
        if do_flush: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    gpu = sum(a.elapsed_time(b) for a, b in evs) / n
    p = _lib.profile_collect(); _lib.profile_enable(False)
    print(f"{label}: gpu {gpu:.3f} ms/step wall {wall:.3f} ms/step", {k: round(v[0] / max(v[1], 1), 3) for k, v in p.items()})
timeit("noprof noflush", False, False)
timeit("noprof flush", False, True)
timeit("prof noflush", True, False)
timeit("prof flush", True, True)
# host-side cost of one step without GPU wait
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host launch time {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms")

import cProfile, pstats
for rep in range(3):
    s0 = torch.cuda.memory_stats()
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): step()
    pr.disable(); torch.cuda.synchronize()
    s1 = torch.cuda.memory_stats()
    print("device allocs", s1["num_device_alloc"] - s0["num_device_alloc"], "device frees", s1["num_device_free"] - s0["num_device_free"], "reserved GB", s1["reserved_bytes.all.current"] / 1e9)
    pstats.Stats(pr).sort_stats("tottime").print_stats(8)
