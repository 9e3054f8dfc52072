#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q --timeout 600 > gpurun_out/r2_train_tests.log 2>&1
echo "train tests exit $?"; tail -n 12 gpurun_out/r2_train_tests.log
timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_be.json 2> gpurun_out/bench_r2_train_be.err
DEPGAN_NO_BATCHED_EVAL=1 timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_nobe.json 2> gpurun_out/bench_r2_train_nobe.err
python - <<'PY'
import json
for f in ("bench_r2_train_be", "bench_r2_train_nobe"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["last_losses"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
