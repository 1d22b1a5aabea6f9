#!/usr/bin/env python
"""Per-kernel device time of the dynamic-stage optimizer step (bench.py's train_step workload), eager launches,
via torch.profiler (CUPTI).  Usage on the GPU box: python scripts/profile_train_step.py > gpurun_out/train_prof.txt"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    holder = {}
    orig = torch.cuda.synchronize

    # reuse bench.train_step_ms's construction, but profile 3 eager steps instead of timing
    from torch.profiler import ProfilerActivity, profile
    import dreammesh4d_b200.trainstep as TS
    real_call = TS.DynamicStageStep.__call__
    calls = {"n": 0}

    def wrapped(self, batches, step=0):
        calls["n"] += 1
        if calls["n"] == 6:          # after the sizing + warm-up calls
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                for i in range(3):
                    out = real_call(self, batches, step)
                torch.cuda.synchronize()
            holder["prof"] = prof
            return out
        return real_call(self, batches, step)

    TS.DynamicStageStep.__call__ = wrapped
    res = bench.train_step_ms(dev, None, 5, bench.N_FACES, bench.VIEWS, 32, "C3")
    print({k: v for k, v in res.items() if k != "what"})
    prof = holder["prof"]
    rows = [(e.key, e.device_time_total / 3e3, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"device time per step (sum over kernels): {tot:.3f} ms; kernels per step: {sum(r[2] for r in rows)}")
    for k, ms, n in rows[:45]:
        print(f"{ms:8.3f} ms  x{n:<4d} {k[:110]}")


if __name__ == "__main__":
    main()
