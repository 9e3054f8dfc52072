#!/bin/bash
# round-2 re-entry check: row-kernel tests + kbench, bench A/B (row kernel on / off), then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
export DEPGAN_DEBUG_SYNC=1
timeout 600 python -m pytest tests/test_gpu_conv_row.py -m gpu -x -q --timeout 300 > gpurun_out/r2_row_tests.log 2>&1
echo "row tests exit $?"; tail -n 8 gpurun_out/r2_row_tests.log
unset DEPGAN_DEBUG_SYNC
timeout 300 python scripts/kbench.py tc_3x3_32to32 tc_3x3_96to32 > gpurun_out/r2_kbench_row3.txt 2>&1
cat gpurun_out/r2_kbench_row3.txt
DEPGAN_NO_ROW=1 timeout 300 python scripts/kbench.py tc_3x3_32to32 tc_3x3_96to32 > gpurun_out/r2_kbench_norow3.txt 2>&1
cat gpurun_out/r2_kbench_norow3.txt
timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_row.json 2> gpurun_out/bench_r2_row.err
DEPGAN_NO_ROW=1 timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_norow.json 2> gpurun_out/bench_r2_norow.err
python - <<'PY'
import json
for f in ("bench_r2_row", "bench_r2_norow"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_classes_ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 1800 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2_gpu_tests_all.log 2>&1
echo "gpu tests exit $?"; tail -n 15 gpurun_out/r2_gpu_tests_all.log
