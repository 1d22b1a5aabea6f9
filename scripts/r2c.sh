#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python scripts/profile_sds.py 8 2>&1 | tail -90) > gpurun_out/r2c_sds_profile.txt
(timeout 300 python scripts/profile_dropin.py 2>&1 | tail -70) > gpurun_out/r2c_dropin_profile.txt
head -5 gpurun_out/r2c_dropin_profile.txt
