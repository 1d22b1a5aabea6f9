#!/bin/bash
# Kernel-tuning A/B without rebuilding on the GPU box: build variants of libdm4d.so HERE (nvcc cross-compiles),
#   python scripts/build_variants.py "name=-DFLAG=.. -DFLAG2=.." ...
# ship them with the snapshot, and time each with the same bench on the box (8 views and, for the latency-bound
# single-view regime of the drop-in / strong-scaling paths, 1 view):
#   gpurun -- 'bash scripts/tune_variants.sh > gpurun_out/tune.log 2>&1'
for lib in dreammesh4d_b200/lib/libdm4d.so dreammesh4d_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  for v in ${DM4D_TUNE_VIEWS:-8 1}; do
    echo -n "$(basename $lib) views=$v | "
    DM4D_VIEWS=$v DM4D_LIB_PATH="$PWD/$lib" timeout 200 python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | tail -1
  done
done
