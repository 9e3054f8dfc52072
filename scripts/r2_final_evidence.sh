#!/bin/bash
# final round-2 evidence on the committed build: ncu launch lists (inference forward at batch 64, one training iteration at
# batch 32), --set full on every convolution launch of the inference forward, launch list of the f16x3 forward
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
DEPGAN_ONLY=f16x3 python scripts/split_rate.py 64 > gpurun_out/split_plain.log 2>&1 &&
DEPGAN_ONLY=f16x3 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_split_launches_b64_raw.csv python scripts/split_rate.py 64 > gpurun_out/split_ncu.log 2>&1
ls -la gpurun_out | tail -6
