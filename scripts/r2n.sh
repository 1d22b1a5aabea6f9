#!/bin/bash
mkdir -p gpurun_out
for lib in dreammesh4d_b200/lib/libdm4d.so dreammesh4d_b200/lib/variants/seg1024.so; do
  echo "== $lib"; DM4D_LIB_PATH=$PWD/$lib timeout 300 python scripts/grad_errors.py 2>&1 | tail -5
done > gpurun_out/r2n_grad_errors.txt
cat gpurun_out/r2n_grad_errors.txt
