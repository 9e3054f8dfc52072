#!/bin/bash
# HEAD record, part 3 (one GPU): ncu --set full on the backward-side kernels of one DEP-GAN training iteration at batch 32
# (weight gradients, the critics' 5x5 row kernels, data-gradient / JVP instantiations of the tile kernel, pool backward,
# channel sums): the first 24 matching launches of the profiled iteration
mkdir -p gpurun_out
timeout 230 ncu --profile-from-start off --set full --clock-control none --import-source on -c 24 \
    -k regex:"wgrad_tc_kernel|conv_rowg_kernel<5|maxpool_bwd|channel_sum|conv_tc_kernel<3, [24], [01], [24]" \
    -o gpurun_out/r02_head_train_full -f python scripts/train_iter.py 32 > gpurun_out/r2_head_train_ncu_full.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r02_head_train_full.ncu-rep --page raw --csv > gpurun_out/r02_head_train_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_head_train_full.ncu-rep
wc -l gpurun_out/r02_head_train_full_raw.csv; tail -3 gpurun_out/r2_head_train_ncu_full.log
