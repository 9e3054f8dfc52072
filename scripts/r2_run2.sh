#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_row.json 2> gpurun_out/bench_r2_row.err
DEPGAN_NO_ROW=1 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_norow.json 2> gpurun_out/bench_r2_norow.err
python - <<'PY'
import json
for f in ("bench_r2_row", "bench_r2_norow"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_classes_ms_per_step"])
PY
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/r2_gpu_tests_all.log 2>&1
echo "gpu tests exit $?"; tail -n 15 gpurun_out/r2_gpu_tests_all.log
