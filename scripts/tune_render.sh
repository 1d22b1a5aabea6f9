#!/bin/bash
# Rebuilds libdm4d.so on the GPU box with different render-kernel constants and prints per-kernel times.
for cfg in "" "-DDM4D_WSTAGES=3" "-DDM4D_WCHUNK=32 -DDM4D_WSTAGES=3" "-DDM4D_WCHUNK=128" "-DDM4D_BWD_MIN_BLOCKS=4" "-DDM4D_BWD_MIN_BLOCKS=3 -DDM4D_WSTAGES=3"; do
  DM4D_NVCC_EXTRA="$cfg" python -m dreammesh4d_b200.build --force > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$cfg |', round(d['ms_per_step'],3), 'ms | fwd', round(k['render_forward_kernel']['ms_per_launch'],3), 'bwd', round(k['render_backward_kernel']['ms_per_launch'],3))"
done
python -m dreammesh4d_b200.build --force > /dev/null 2>&1
