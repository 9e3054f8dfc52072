#!/bin/bash
mkdir -p gpurun_out
C="tc_5x5_16to16_N96 tc_5x5_32to32_N96 tc_3x3_64to64_plain_N64 tc_3x3_64to64_filmA_N64 tc_5x5_16to32 tc_5x5_32to16"
export DEPGAN_DEBUG_SYNC=1
timeout 900 python -m pytest tests/test_gpu_conv_rowg.py -m gpu -x -q --timeout 300 > gpurun_out/r2_rowg_tests.log 2>&1
echo "rowg tests exit $?"; tail -n 5 gpurun_out/r2_rowg_tests.log
unset DEPGAN_DEBUG_SYNC
for v in "16 16" "0 8" "0 16" "16 8" "32 16" "8 16"; do
  set -- $v
  echo "== pf $1 na_cap $2"
  DEPGAN_ROWG_PF=$1 DEPGAN_ROWG_NA=$2 timeout 200 python scripts/kbench.py $C
done > gpurun_out/r2_rowg_pf.txt 2>&1
cat gpurun_out/r2_rowg_pf.txt
