#!/bin/bash
# Round-2 collection run: GPU tests, smoke, the four bench lines, ncu launch list + full captures of the step's kernels.
mkdir -p gpurun_out
T=${1:-r2v}
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -8) > gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
(timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2)
(timeout 700 python bench.py --steps 20 --warmup 3 2> gpurun_out/${T}_bench.err) > gpurun_out/${T}_bench_n1.json
(timeout 300 python bench.py --config c4 --steps 10 --warmup 3 2> gpurun_out/${T}_c4.err) > gpurun_out/${T}_bench_c4.json
(timeout 300 python bench.py --config c5 2> gpurun_out/${T}_c5.err) > gpurun_out/${T}_bench_c5.json
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 1500) > gpurun_out/${T}_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --kernels-only > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"render_|sort_pack|preprocess|scatter|scan_tiles" -s 30 -c 11 -o gpurun_out/${T}_step -f python bench.py --steps 2 --warmup 3 --no-graph --kernels-only > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"skin_" -s 40 -c 10 -o gpurun_out/${T}_skin -f python bench.py --config c5 --steps 2 --warmup 3 --no-graph > gpurun_out/${T}_ncu_skin.log 2>&1
tail -1 gpurun_out/${T}_ncu_skin.log
