#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_edge5.py tests/test_gpu_train.py tests/test_gpu_nets.py tests/test_gpu_conv.py -m gpu -x -q --timeout 600 > gpurun_out/r2_tests15.log 2>&1
echo "tests exit $?"; tail -n 15 gpurun_out/r2_tests15.log
C="first_5x5 last_5x5 wgrad_first_5x5"
echo "== new"; timeout 200 python scripts/kbench.py $C
echo "== DEPGAN_NO_FIRST5=1"; DEPGAN_NO_FIRST5=1 timeout 200 python scripts/kbench.py $C
timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_e5.json 2> gpurun_out/bench_r2_train_e5.err
DEPGAN_NO_FIRST5=1 timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_noe5.json 2> gpurun_out/bench_r2_train_noe5.err
python - <<'PY'
import json
for f in ("bench_r2_train_e5", "bench_r2_train_noe5"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
