#!/bin/bash
# round-2 baseline diagnostics: kbench of the 256x256 layers, role traces, ncu source page of the FiLM / head kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
python scripts/kbench.py first_3x3 tc_3x3_32to32 tc_3x3_96to32 tc_deconv_64_N64 tc_3x3_64to64 tc_3x3_160 > gpurun_out/r2_kbench0.txt 2>&1
DEPGAN_B200_LIB=build_ab/libtrace.so python scripts/trace_tc.py tc_3x3_32to32_plain_N64 tc_3x3_32to32_filmA_N64 tc_3x3_32to32_head_N64 tc_3x3_96to32_N64 > gpurun_out/r2_trace0.txt 2>&1
export KBENCH_REPS=1 KBENCH_WARMUP=0
CASES="tc_3x3_32to32_plain_N64 tc_3x3_32to32_filmA_N64 tc_3x3_32to32_head_N64 tc_3x3_96to32_N64"
python scripts/kbench.py $CASES > gpurun_out/r2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -o gpurun_out/r2_tc0 -f python scripts/kbench.py $CASES > gpurun_out/r2_ncu0.log 2>&1
ls -la gpurun_out | tail -5
