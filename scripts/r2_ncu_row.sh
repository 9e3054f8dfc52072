#!/bin/bash
mkdir -p gpurun_out
export KBENCH_REPS=1 KBENCH_WARMUP=0
CASES="tc_3x3_32to32_plain_N64 tc_3x3_32to32_filmA_N64 tc_3x3_32to32_head_N64 tc_3x3_96to32_N64"
python scripts/kbench.py $CASES > gpurun_out/r2_plain_row.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_row_kernel" -o gpurun_out/r2_row1 -f python scripts/kbench.py $CASES > gpurun_out/r2_ncu_row1.log 2>&1
ls -la gpurun_out | tail -3
