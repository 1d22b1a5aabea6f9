#!/bin/bash
for cfg in "" "-DDM4D_SORT_THREADS=1024 -DDM4D_SORT_CHUNK=8192" "-DDM4D_SORT_THREADS=1024" "-DDM4D_SORT_CHUNK=8192" "-DDM4D_SORT_THREADS=1024 -DDM4D_SORT_CHUNK=16384" "-DDM4D_SORT_THREADS=256"; do
  DM4D_NVCC_EXTRA="$cfg" python -m dreammesh4d_b200.build --force > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  timeout 200 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$cfg |', round(d['ms_per_step'],3), 'ms | sort', round(k['sort_pack_kernel']['ms_per_launch'],3))"
done
python -m dreammesh4d_b200.build --force > /dev/null 2>&1
