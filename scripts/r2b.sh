#!/bin/bash
# round 2, call b: validate the new paths (SDS step, plan re-use, capacity book, exchange modes) and the new bench legs
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "^E  *[+|]" | tail -25) > gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
(timeout 700 python bench.py --steps 20 --warmup 3 2> gpurun_out/r2b_bench.err) > gpurun_out/r2b_bench.json
tail -25 gpurun_out/r2b_bench.err
(timeout 200 python bench.py --config c4 --steps 10 --warmup 3 2> gpurun_out/r2b_c4.err) > gpurun_out/r2b_c4.json
(timeout 200 python bench.py --config c5 --steps 20 --warmup 3 2> gpurun_out/r2b_c5.err) > gpurun_out/r2b_c5.json
tail -3 gpurun_out/r2b_c4.err gpurun_out/r2b_c5.err
head -c 1500 gpurun_out/r2b_c5.json
