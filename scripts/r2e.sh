#!/bin/bash
# 2-GPU call: the N=2 bench line (weak + strong legs, exchange in the graph, train-step exchange modes) and the new tests
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_nhwc_gpu.py -q -m gpu -x 2>&1 | tail -4) > gpurun_out/r2e_tests.log
tail -3 gpurun_out/r2e_tests.log
(timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/r2e_bench_n2.err) > gpurun_out/r2e_bench_n2.json
grep -E "bench|Error|error" gpurun_out/r2e_bench_n2.err | tail -25
