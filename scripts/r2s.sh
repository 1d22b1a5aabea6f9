#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_properties.py tests/test_deform_graph.py tests/test_plugin.py -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -12) > gpurun_out/r2s_tests.log
tail -6 gpurun_out/r2s_tests.log
# compute-sanitizer on the small parity scenes: memcheck, then racecheck (shared-memory hazards of the per-warp rings)
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_raster_parity_gpu.py -q -m gpu -k "dropin_single_view or empty_and_fully or capacity_mode" 2>&1 | tail -15) > gpurun_out/r2s_memcheck.log
tail -6 gpurun_out/r2s_memcheck.log
(timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_raster_parity_gpu.py -q -m gpu -k "dropin_single_view and 64" 2>&1 | tail -15) > gpurun_out/r2s_racecheck.log
tail -6 gpurun_out/r2s_racecheck.log
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_skin_parity_gpu.py tests/test_nhwc_gpu.py -q -m gpu -k "golden or fp16" 2>&1 | tail -8) > gpurun_out/r2s_memcheck_skin_norm.log
tail -4 gpurun_out/r2s_memcheck_skin_norm.log
