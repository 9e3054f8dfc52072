#!/bin/bash
# last round-2 record on the committed HEAD (one GPU, ~5 min): GPU suite, smoke(), default bench line, launch list of
# the inference forward at batch 64
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/r2_head_gpu_tests.log 2>&1
echo "gpu tests exit $?"; tail -n 2 gpurun_out/r2_head_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py > gpurun_out/bench_r2_head_1gpu.json 2> gpurun_out/bench_r2_head_1gpu.err
echo "bench exit $?"
timeout 40 python scripts/infer_iter.py 64 > gpurun_out/r2_head_infer_plain.log 2>&1 &&
timeout 90 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_head_infer_launches_b64_raw.csv python scripts/infer_iter.py 64 > gpurun_out/r2_head_infer_ncu.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_head_1gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"])
print("train", d["train"]["value"], d["train"]["ms_per_step"])
PY
