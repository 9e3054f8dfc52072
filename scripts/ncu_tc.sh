#!/bin/bash
# ncu --set full on a few conv_tc_kernel launches of scripts/kbench.py; only small CSV exports are kept.
mkdir -p gpurun_out
export KBENCH_REPS=1 KBENCH_WARMUP=0
CASES="${@:-tc_3x3_32to32_plain_N64 tc_3x3_32to32_film_N64 tc_3x3_64to64_film_N64 tc_deconv_64_N64}"
ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel" -o /tmp/tc -f python scripts/kbench.py $CASES > gpurun_out/ncu_tc.log 2>&1
ncu -i /tmp/tc.ncu-rep --page raw --csv > gpurun_out/ncu_tc_raw.csv 2>/dev/null
ncu -i /tmp/tc.ncu-rep --page details --csv > gpurun_out/ncu_tc_details.csv 2>/dev/null
ncu -i /tmp/tc.ncu-rep --page source --csv --kernel-name regex:conv_tc_kernel --launch-skip 1 --launch-count 1 > gpurun_out/ncu_tc_source_film.csv 2>/dev/null
ls -la gpurun_out /tmp/tc.ncu-rep
