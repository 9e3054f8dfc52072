#!/bin/bash
# HEAD record, part 2 (one GPU, ~3 min): ncu launch list of one DEP-GAN training iteration at batch 32, reference arm
mkdir -p gpurun_out
timeout 60 python scripts/train_iter.py 32 > gpurun_out/r2_head_train_plain.log 2>&1 &&
timeout 120 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_head_train_launches_b32_raw.csv python scripts/train_iter.py 32 > gpurun_out/r2_head_train_ncu.log 2>&1
echo "ncu exit $?"
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_head_ref.json 2> gpurun_out/bench_r2_head_ref.err
echo "ref exit $?"; tail -c 400 gpurun_out/bench_r2_head_ref.json
