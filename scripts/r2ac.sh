#!/bin/bash
# Round 2, last session: ncu evidence of the FINAL tree (launch list + full captures of the step kernels, the skinning
# kernels and the streamed GroupNorm kernels) and the plain GroupNorm timing.
mkdir -p gpurun_out
T=r2ac
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --kernels-only > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"render_|sort_pack|preprocess|scatter|scan_tiles" -s 33 -c 11 -o gpurun_out/${T}_step -f python bench.py --steps 2 --warmup 3 --no-graph --kernels-only > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_stream" -s 8 -c 4 -o gpurun_out/${T}_gn -f python scripts/bench_gn.py > gpurun_out/${T}_ncu_gn.log 2>&1
tail -1 gpurun_out/${T}_ncu_gn.log
timeout 120 python scripts/bench_gn.py > gpurun_out/${T}_groupnorm.txt 2>&1
cat gpurun_out/${T}_groupnorm.txt
