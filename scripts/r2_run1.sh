#!/bin/bash
mkdir -p gpurun_out
export DEPGAN_TEST_LOG=gpurun_out/r2_test_values.jsonl
rm -f $DEPGAN_TEST_LOG
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q --timeout 600 > gpurun_out/r2_test_bench_shapes.log 2>&1
echo "tests exit $?"; tail -n 30 gpurun_out/r2_test_bench_shapes.log
( time timeout 900 python bench.py ) > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r2_a.err
