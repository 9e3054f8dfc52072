#!/bin/bash
# ncu evidence for the inference step: (1) launch list of one batch-64 forward, (2) --set full on the conv_tc_kernel
# launches of one batch-16 forward (DRAM traffic, tensor-pipe activity).  Only small CSV exports are kept.
set -e
mkdir -p gpurun_out
python scripts/infer_iter.py 64 > gpurun_out/infer_iter_plain.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/infer_launches_b64.csv python scripts/infer_iter.py 64 > gpurun_out/infer_iter_ncu1.log 2>&1
python scripts/infer_iter.py 16 > gpurun_out/infer_iter_plain16.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc_kernel -o /tmp/tcfull -f \
    python scripts/infer_iter.py 16 > gpurun_out/infer_iter_ncu2.log 2>&1
ncu -i /tmp/tcfull.ncu-rep --page raw --csv > gpurun_out/infer_conv_tc_full_raw_b16.csv 2>/dev/null
ls -la gpurun_out /tmp/tcfull.ncu-rep
