#!/bin/bash
# round-2 evidence run: ncu launch lists (inference forward, training iteration) and --set full on the convolution kernels of
# the inference forward + the row-kernel kbench cases, all on the committed build
mkdir -p gpurun_out
python scripts/infer_iter.py 64 > gpurun_out/infer_iter_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_infer_launches_b64_raw.csv python scripts/infer_iter.py 64 > gpurun_out/infer_iter_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_row_kernel|conv_rowg_kernel" -o gpurun_out/r02_conv_full -f \
    python scripts/infer_iter.py 64 > gpurun_out/infer_iter_ncu2.log 2>&1
ncu -i gpurun_out/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_full_raw_b64.csv 2>/dev/null
rm -f gpurun_out/r02_conv_full.ncu-rep
python scripts/train_iter.py 32 > gpurun_out/train_iter_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_train_launches_b32_raw.csv python scripts/train_iter.py 32 > gpurun_out/train_iter_ncu.log 2>&1
export KBENCH_REPS=1 KBENCH_WARMUP=1
CASES="tc_5x5_16to16_N96 tc_5x5_32to32_N96 tc_5x5_16to32_N96 tc_5x5_32to16_mask_N96 tc_3x3_64to64_plain_N64 tc_3x3_64to64_filmA_N64 tc_3x3_32to32_pool_N64"
python scripts/kbench.py $CASES > gpurun_out/r2_kb_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"conv_rowg_kernel" -o gpurun_out/r02_rowg_full -f python scripts/kbench.py $CASES > gpurun_out/r2_ncu_rowg.log 2>&1
ncu -i gpurun_out/r02_rowg_full.ncu-rep --page raw --csv > gpurun_out/r02_rowg_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_rowg_full.ncu-rep
ls -la gpurun_out | tail -8
