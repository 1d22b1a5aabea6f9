#!/bin/bash
# usage: scripts/gpu_round.sh <tag> : GPU tests + bench lines, outputs under gpurun_out/<tag>_*
tag=$1
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --config c5 > gpurun_out/${tag}_c5.json 2> gpurun_out/${tag}_c5.err
python bench.py --config c4 > gpurun_out/${tag}_c4.json 2> gpurun_out/${tag}_c4.err
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python - <<P
import json
for f in ("c5","c4","bench"):
    try:
        d=json.loads(open(f"gpurun_out/${tag}_{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    if f=="c5":
        print("c5", {k:(v["fwd_bwd_us"],v["hbm_frac"],v["kernels_us"]) for k,v in d["methods"].items()}, d.get("batched_timestamps"))
    else:
        print(f, d["ms_per_step"], {k:round(v["ms_per_launch"],4) for k,v in d["kernels"].items()}, d.get("dropin",{}).get("ratio_to_batched"), d.get("train_step",{}).get("ms_graph"))
P
