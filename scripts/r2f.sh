#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -40) > gpurun_out/r2f_tests.log
tail -12 gpurun_out/r2f_tests.log
for v in 1 8; do echo -n "views=$v | "; DM4D_VIEWS=$v timeout 200 python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | tail -1; done > gpurun_out/r2f_views.log
cat gpurun_out/r2f_views.log
