#!/bin/bash
mkdir -p gpurun_out
export DEPGAN_DEBUG_SYNC=1
timeout 900 python -m pytest tests/test_gpu_conv_rowg.py -m gpu -x -q --timeout 300 > gpurun_out/r2_rowg_tests.log 2>&1
rc=$?; echo "rowg tests exit $rc"; tail -n 25 gpurun_out/r2_rowg_tests.log
unset DEPGAN_DEBUG_SYNC
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python scripts/kbench.py tc_5x5 tc_3x3_64to64 tc_3x3_32to64 > gpurun_out/r2_kbench_rowg5.txt 2>&1; cat gpurun_out/r2_kbench_rowg5.txt
DEPGAN_B200_LIB=build_ab/librowg4.so timeout 300 python scripts/kbench.py tc_5x5 tc_3x3_64to64 tc_3x3_32to64 > gpurun_out/r2_kbench_rowg4b.txt 2>&1; cat gpurun_out/r2_kbench_rowg4b.txt
