#!/bin/bash
# A/B of libdm4d.so variants on the C5 skinning microbench (per-kernel us at 8 timestamps)
for lib in dreammesh4d_b200/lib/libdm4d.so dreammesh4d_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  echo -n "$(basename $lib) | "
  DM4D_LIB_PATH="$PWD/$lib" timeout 200 python bench.py --config c5 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['batched_timestamps']['hybrid']; print(round(d['methods']['hybrid']['fwd_bwd_us'],1), round(b['fwd_bwd_us'],1), b['kernels_us'])"
done
