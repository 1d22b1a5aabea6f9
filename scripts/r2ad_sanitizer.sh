#!/bin/bash
# compute-sanitizer on the kernels of the last session of round 2: streamed GroupNorm (statistics / apply, forward and
# backward, fp32 and fp16) and the 256 x 8 sort tier at its boundaries.
mkdir -p gpurun_out
(timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_nhwc_gpu.py -q -m gpu -k "groupnorm" 2>&1 | tail -8) > gpurun_out/r2ad_memcheck_groupnorm.log
tail -3 gpurun_out/r2ad_memcheck_groupnorm.log
(timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_raster_parity_gpu.py -q -m gpu -k "tier_boundaries and (1025 or 2048 or 2049 or 4097)" 2>&1 | tail -8) > gpurun_out/r2ad_memcheck_sort_tiers.log
tail -3 gpurun_out/r2ad_memcheck_sort_tiers.log
(timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_nhwc_gpu.py tests/test_raster_parity_gpu.py -q -m gpu -k "(groupnorm and fp16) or (groupnorm and 128 and True) or (tier_boundaries and 2049)" 2>&1 | tail -8) > gpurun_out/r2ad_racecheck_groupnorm_sort.log
tail -3 gpurun_out/r2ad_racecheck_groupnorm_sort.log
