#!/bin/bash
# compute-sanitizer on the kernels added in the second session of round 2 (small parity scenes)
mkdir -p gpurun_out
(timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_raster_parity_gpu.py -q -m gpu -k "tier_boundaries and (1025 or 8193 or 16385 or 21000)" 2>&1 | tail -8) > gpurun_out/r2w_memcheck_sort_tiers.log
tail -3 gpurun_out/r2w_memcheck_sort_tiers.log
(timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_skin_parity_gpu.py -q -m gpu -k "ragged or incidence" 2>&1 | tail -8) > gpurun_out/r2w_memcheck_skin.log
tail -3 gpurun_out/r2w_memcheck_skin.log
(timeout 800 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_nhwc_gpu.py -q -m gpu -k "add_layernorm or geglu or bias_residual or (groupnorm and 320)" 2>&1 | tail -8) > gpurun_out/r2w_memcheck_nhwc.log
tail -3 gpurun_out/r2w_memcheck_nhwc.log
(timeout 800 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_raster_parity_gpu.py tests/test_skin_parity_gpu.py -q -m gpu -k "(tier_boundaries and 1025) or (ragged and 1002)" 2>&1 | tail -8) > gpurun_out/r2w_racecheck_sort_skin.log
tail -3 gpurun_out/r2w_racecheck_sort_skin.log
