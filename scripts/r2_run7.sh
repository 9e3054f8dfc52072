#!/bin/bash
# round-2: generic row kernel v2 (phantom slots, lean issuer) -- parity, kbench, role trace
mkdir -p gpurun_out
export DEPGAN_DEBUG_SYNC=1
timeout 900 python -m pytest tests/test_gpu_conv_rowg.py -m gpu -x -q --timeout 300 > gpurun_out/r2_rowg_tests.log 2>&1
rc=$?; echo "rowg tests exit $rc"; tail -n 25 gpurun_out/r2_rowg_tests.log
unset DEPGAN_DEBUG_SYNC
if [ $rc -ne 0 ]; then exit 0; fi
CASES="tc_5x5 tc_3x3_64to64 tc_3x3_32to64"
timeout 300 python scripts/kbench.py $CASES > gpurun_out/r2_kbench_rowg2.txt 2>&1; cat gpurun_out/r2_kbench_rowg2.txt
TRACE_ROWG=1 DEPGAN_B200_LIB=build_ab/librowgtrace.so python scripts/trace_row.py tc_5x5_16to16_N96 tc_5x5_32to32_N96 tc_3x3_64to64_plain_N64 tc_3x3_64to64_filmA_N64 > gpurun_out/r2_trace_rowg2.txt 2>&1
timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_rowg2.json 2> gpurun_out/bench_r2_rowg2.err
timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_rowg2.json 2> gpurun_out/bench_r2_train_rowg2.err
python - <<'PY'
import json
for f in ("bench_r2_rowg2",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["dtype"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_classes_ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
for f in ("bench_r2_train_rowg2",):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
