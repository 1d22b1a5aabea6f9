#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_backward -s 3 -c 1 -o gpurun_out/r2j_bwd python bench.py --steps 2 --warmup 3 --no-graph --kernels-only > gpurun_out/r2j_ncu.log 2>&1
tail -3 gpurun_out/r2j_ncu.log
ls -la gpurun_out/r2j_bwd.ncu-rep
