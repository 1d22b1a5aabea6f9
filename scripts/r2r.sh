#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v "^E  *[+|]" | tail -30) > gpurun_out/r2r_tests.log
tail -15 gpurun_out/r2r_tests.log
(timeout 200 python bench.py --config c5 --steps 20 --warmup 3 2> gpurun_out/r2r_c5.err) > gpurun_out/r2r_c5.json
python -c "
import json; j=json.load(open('gpurun_out/r2r_c5.json'))
for m,v in j['methods'].items(): print(m, round(v['fwd_bwd_us'],1),'us', round(v['GBps'],1),'GB/s', v['hbm_frac'], v['kernels_us'])
"
