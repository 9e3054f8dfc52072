#!/bin/bash
# round-2: generic row kernel (conv_rowg.cu) -- parity, kbench A/B against the tile kernel, inference bench A/B, train bench
mkdir -p gpurun_out
export DEPGAN_DEBUG_SYNC=1
timeout 900 python -m pytest tests/test_gpu_conv_rowg.py -m gpu -x -q --timeout 300 > gpurun_out/r2_rowg_tests.log 2>&1
rc=$?; echo "rowg tests exit $rc"; tail -n 25 gpurun_out/r2_rowg_tests.log
unset DEPGAN_DEBUG_SYNC
if [ $rc -ne 0 ]; then exit 0; fi
CASES="tc_5x5 tc_3x3_64to64 tc_3x3_32to64"
timeout 300 python scripts/kbench.py $CASES > gpurun_out/r2_kbench_rowg.txt 2>&1; cat gpurun_out/r2_kbench_rowg.txt
DEPGAN_NO_ROWG=1 timeout 300 python scripts/kbench.py $CASES > gpurun_out/r2_kbench_norowg.txt 2>&1; cat gpurun_out/r2_kbench_norowg.txt
timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_rowg.json 2> gpurun_out/bench_r2_rowg.err
DEPGAN_NO_ROWG=1 timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_norowg.json 2> gpurun_out/bench_r2_norowg.err
timeout 600 python bench.py --no-train --no-extra --no-cpu --precision bf16 > gpurun_out/bench_r2_rowg_bf16.json 2> gpurun_out/bench_r2_rowg_bf16.err
python - <<'PY'
import json
for f in ("bench_r2_rowg", "bench_r2_norowg", "bench_r2_rowg_bf16"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["dtype"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_classes_ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_rowg.json 2> gpurun_out/bench_r2_train_rowg.err
DEPGAN_NO_ROWG=1 timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_norowg.json 2> gpurun_out/bench_r2_train_norowg.err
python - <<'PY'
import json
for f in ("bench_r2_train_rowg", "bench_r2_train_norowg"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("conv_time_per_iteration"))
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 600 --deselect tests/test_gpu_conv_rowg.py > gpurun_out/r2_gpu_tests_all.log 2>&1
echo "gpu tests exit $?"; tail -n 8 gpurun_out/r2_gpu_tests_all.log
