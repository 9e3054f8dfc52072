#!/bin/bash
mkdir -p gpurun_out
DEPGAN_B200_LIB=build_ab/librowtrace.so python scripts/trace_row.py tc_3x3_32to32_plain_N64 tc_3x3_96to32_N64 tc_3x3_32to32_filmA_N64 > gpurun_out/r2_trace_row.txt 2>&1
tail -5 gpurun_out/r2_trace_row.txt
