#!/bin/bash
# First GPU call of the next round: re-validate the tree and run the experiments prepared at the end of round 1.
#   python scripts/build_variants.py "cells4x2=-DDM4D_CELL_ROWS=2" "fastexp=-DDM4D_FAST_EXP"   # here, before the call
#   gpurun --timeout 900 -- 'bash scripts/next_round_first_call.sh'
mkdir -p gpurun_out
(timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2) > gpurun_out/nr_smoke.log
(timeout 400 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -15) > gpurun_out/nr_tests.log
(timeout 300 python bench.py --steps 20 --warmup 3 2> gpurun_out/nr_bench.err) > gpurun_out/nr_bench.json
(timeout 120 python scripts/exp_split_streams.py 2 2>&1 | tail -4) > gpurun_out/nr_split2.log
(timeout 120 python scripts/exp_split_streams.py 4 2>&1 | tail -4) > gpurun_out/nr_split4.log
bash scripts/tune_variants.sh > gpurun_out/nr_tune.log 2>&1
# parity of the experimental 4x2-cell build (never run on a GPU in round 1)
if [ -f dreammesh4d_b200/lib/variants/cells4x2.so ]; then
  (DM4D_LIB_PATH=$PWD/dreammesh4d_b200/lib/variants/cells4x2.so timeout 300 python -m pytest tests/test_raster_parity_gpu.py tests/test_full_size_gpu.py -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -12) > gpurun_out/nr_cells4x2_tests.log
  tail -4 gpurun_out/nr_cells4x2_tests.log
fi
(timeout 200 python scripts/profile_train_step.py 2>&1 | head -40) > gpurun_out/nr_train_prof.txt
cat gpurun_out/nr_smoke.log; tail -4 gpurun_out/nr_tests.log; cat gpurun_out/nr_split2.log gpurun_out/nr_split4.log gpurun_out/nr_tune.log
