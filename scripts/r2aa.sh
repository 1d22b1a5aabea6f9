#!/bin/bash
# Round 2, last session: finer sort tiers (256 x 8 for 1025..2048 keys, 64 x 8 for <= 512 keys) at C3 and C4.
mkdir -p gpurun_out
out=gpurun_out/r2aa_tune_sort_tiers.txt
: > $out
for lib in dreammesh4d_b200/lib/libdm4d.so dreammesh4d_b200/lib/variants/*.so; do
  echo -n "$(basename $lib) C3 views=8 | " >> $out
  DM4D_VIEWS=8 DM4D_LIB_PATH="$PWD/$lib" timeout 200 python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | tail -1 >> $out
done
for lib in dreammesh4d_b200/lib/variants/t256t64.so; do
  echo -n "$(basename $lib) C3 views=1 | " >> $out
  DM4D_VIEWS=1 DM4D_LIB_PATH="$PWD/$lib" timeout 200 python bench.py --steps 10 --warmup 3 --kernels-only 2>/dev/null | tail -1 >> $out
done
for lib in dreammesh4d_b200/lib/libdm4d.so dreammesh4d_b200/lib/variants/t256.so dreammesh4d_b200/lib/variants/t256t64.so; do
  echo -n "$(basename $lib) C4 | " >> $out
  DM4D_LIB_PATH="$PWD/$lib" timeout 300 python bench.py --config c4 --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'ms_per_step': d['ms_per_step'], 'kernels_ms': {k: round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}}))" >> $out
done
cat $out
for lib in dreammesh4d_b200/lib/variants/t256t64.so; do
  echo "== $(basename $lib): tier-boundary + full-size parity" >> $out
  DM4D_LIB_PATH="$PWD/$lib" timeout 400 python -m pytest tests/test_raster_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -3 >> $out
done
tail -4 $out
