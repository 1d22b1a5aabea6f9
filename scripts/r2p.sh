#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -12) > gpurun_out/r2p_tests.log
tail -5 gpurun_out/r2p_tests.log
(timeout 100 python __graft_entry__.py smoke 2>&1 | tail -2)
(timeout 700 python bench.py --steps 20 --warmup 3 2> gpurun_out/r2p_bench.err) > gpurun_out/r2p_bench.json
grep bench gpurun_out/r2p_bench.err | tail -14
