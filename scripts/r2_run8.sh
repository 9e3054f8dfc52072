#!/bin/bash
# round-2: launch list of one training iteration (ncu, every kernel) + kbench of the v4 row kernel
mkdir -p gpurun_out
timeout 300 python scripts/kbench.py tc_5x5 tc_3x3_64to64 tc_3x3_32to64 > gpurun_out/r2_kbench_rowg4.txt 2>&1; cat gpurun_out/r2_kbench_rowg4.txt
python scripts/train_iter.py 32 > gpurun_out/train_iter_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_train_launches_b32_raw.csv python scripts/train_iter.py 32 > gpurun_out/train_iter_ncu.log 2>&1
ls -la gpurun_out/r02_train_launches_b32_raw.csv
timeout 900 python bench.py --workload depgan_train --steps 10 --warmup 3 > gpurun_out/bench_r2_train_rowg4.json 2> gpurun_out/bench_r2_train_rowg4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_train_rowg4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
PY
