#!/bin/bash
# round-2: f16 inference path -- tests, then bench A/B (f16 vs bf16), then the remaining GPU suite
mkdir -p gpurun_out
export DEPGAN_TEST_LOG=$PWD/gpurun_out/r2_test_values.jsonl
rm -f $DEPGAN_TEST_LOG
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_c_caller.py tests/test_gpu_bench_shapes.py -m gpu -x -q --timeout 600 > gpurun_out/r2_f16_tests.log 2>&1
echo "f16 tests exit $?"; tail -n 12 gpurun_out/r2_f16_tests.log
grep fresh_init $DEPGAN_TEST_LOG
timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_f16.json 2> gpurun_out/bench_r2_f16.err
timeout 600 python bench.py --no-train --no-extra --no-cpu --precision bf16 > gpurun_out/bench_r2_bf16.json 2> gpurun_out/bench_r2_bf16.err
python - <<'PY'
import json
for f in ("bench_r2_f16", "bench_r2_bf16"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["dtype"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_classes_ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 600 --deselect tests/test_gpu_nets.py --deselect tests/test_c_caller.py --deselect tests/test_gpu_bench_shapes.py > gpurun_out/r2_gpu_tests_rest.log 2>&1
echo "rest of gpu tests exit $?"; tail -n 6 gpurun_out/r2_gpu_tests_rest.log
