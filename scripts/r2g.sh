#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -40) > gpurun_out/r2g_tests.log
tail -12 gpurun_out/r2g_tests.log
bash scripts/tune_variants.sh > gpurun_out/r2g_tune.log 2>&1
cat gpurun_out/r2g_tune.log
