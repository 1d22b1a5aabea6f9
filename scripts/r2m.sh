#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_full_size_gpu.py -q -m gpu -k c3 2>&1 | grep -E "assert|Error|rel|view|^E " | head -30) > gpurun_out/r2m_tests.log
cat gpurun_out/r2m_tests.log
(DM4D_LIB_PATH=$PWD/dreammesh4d_b200/lib/variants/seg1024.so timeout 600 python -m pytest tests/test_full_size_gpu.py -q -m gpu -k c3 2>&1 | tail -3)
