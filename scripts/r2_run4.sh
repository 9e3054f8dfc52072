#!/bin/bash
# round-2: whole GPU suite, MMA-rate micro-benchmark, ncu evidence of the inference step (launch list + --set full), full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
export DEPGAN_TEST_LOG=$PWD/gpurun_out/r2_test_values.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2_gpu_tests_all.log 2>&1
echo "gpu tests exit $?"; tail -n 6 gpurun_out/r2_gpu_tests_all.log
unset DEPGAN_TEST_LOG
scripts/ubench/mma_rate > gpurun_out/r2_mma_rate.txt 2>&1; cat gpurun_out/r2_mma_rate.txt
python scripts/infer_iter.py 64 > gpurun_out/infer_iter_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_infer_launches_b64_raw.csv python scripts/infer_iter.py 64 > gpurun_out/infer_iter_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv_row_kernel" -o gpurun_out/r02_conv_full -f \
    python scripts/infer_iter.py 64 > gpurun_out/infer_iter_ncu2.log 2>&1
ncu -i gpurun_out/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_full_raw_b64.csv 2>/dev/null
timeout 1500 python bench.py > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_r2_a.json
ls -la gpurun_out | tail -12
