#!/bin/bash
mkdir -p gpurun_out
TRACE_ROWG=1 DEPGAN_B200_LIB=build_ab/librowgtrace.so python scripts/trace_row.py tc_5x5_16to16_N96 tc_5x5_32to32_N96 tc_3x3_64to64_plain_N64 tc_3x3_64to64_filmA_N64 > gpurun_out/r2_trace_rowg.txt 2>&1
tail -3 gpurun_out/r2_trace_rowg.txt
