#!/bin/bash
# ncu --set full on the weight-gradient and first-layer kernels (scripts/kbench.py cases); small CSV export only.
mkdir -p gpurun_out
export KBENCH_REPS=1 KBENCH_WARMUP=1
python scripts/kbench.py wgrad first_5x5 > gpurun_out/kbench_plain_wg.txt 2>&1 || exit 1
ncu --set full --clock-control none -k regex:"wgrad_tc_kernel|wgrad_first_tc_kernel|conv_first_tc_kernel" -o /tmp/wg -f \
    python scripts/kbench.py wgrad first_5x5 > gpurun_out/ncu_wg.log 2>&1
ncu -i /tmp/wg.ncu-rep --page raw --csv > gpurun_out/ncu_wg_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_wg_raw.csv
