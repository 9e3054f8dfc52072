#!/bin/bash
mkdir -p gpurun_out
export DEPGAN_DEBUG_SYNC=1
timeout 600 python -m pytest tests/test_gpu_conv_row.py -m gpu -x -q --timeout 300 > gpurun_out/r2_row_tests.log 2>&1
echo "row tests exit $?"; tail -n 30 gpurun_out/r2_row_tests.log
unset DEPGAN_DEBUG_SYNC
timeout 300 python scripts/kbench.py tc_3x3_32to32 tc_3x3_96to32 > gpurun_out/r2_kbench_row2.txt 2>&1
cat gpurun_out/r2_kbench_row2.txt
DEPGAN_B200_LIB=build_ab/librowtrace.so timeout 120 python scripts/trace_row.py tc_3x3_32to32_plain_N64 tc_3x3_96to32_N64 tc_3x3_32to32_filmA_N64 > gpurun_out/r2_trace_row.txt 2>&1
