#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_raster_parity_gpu.py tests/test_full_size_gpu.py tests/test_renderer_gpu.py -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -8) > gpurun_out/r2k_tests.log
tail -4 gpurun_out/r2k_tests.log
bash scripts/tune_variants.sh > gpurun_out/r2k_tune.log 2>&1
cat gpurun_out/r2k_tune.log
