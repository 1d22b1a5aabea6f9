#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/r2q_bench_n$N.err) > gpurun_out/r2q_bench_n$N.json
grep -E "bench|Error|error|Timeout" gpurun_out/r2q_bench_n$N.err | sort | uniq -c | sort -k3 | tail -20
(timeout 200 python bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | head -c 400)
