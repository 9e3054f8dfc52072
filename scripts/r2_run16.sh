#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_nets.py tests/test_gpu_infer_dp.py tests/test_c_caller.py -m gpu -x -q --timeout 600 > gpurun_out/r2_tests16.log 2>&1
echo "tests exit $?"; tail -n 5 gpurun_out/r2_tests16.log
timeout 900 python bench.py --no-train --no-cpu > gpurun_out/bench_r2_predict.json 2> gpurun_out/bench_r2_predict.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_predict.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
print("predict", d.get("predict_numpy"))
print("cfg0 e2e", d["configs"]["configs[0]"]["e2e"])
print("cohort", d["configs"]["configs[4]"]["value"])
PY
