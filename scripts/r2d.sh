#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_nhwc_gpu.py tests/test_dropin_fastpath_gpu.py -q -m gpu -x 2>&1 | grep -v "^E  *[+|]" | tail -25) > gpurun_out/r2d_tests.log
tail -8 gpurun_out/r2d_tests.log
(timeout 300 python scripts/profile_sds.py 8 2>&1 | tail -90) > gpurun_out/r2d_sds_profile.txt
grep -E "Self CUDA time total|=====" gpurun_out/r2d_sds_profile.txt
for v in 1 2 4 8; do echo -n "views=$v | "; DM4D_VIEWS=$v timeout 200 python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | tail -1; done > gpurun_out/r2d_views.log
cat gpurun_out/r2d_views.log
