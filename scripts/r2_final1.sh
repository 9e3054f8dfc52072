#!/bin/bash
# round-2 final check on one GPU: the whole GPU suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
export DEPGAN_TEST_LOG=$PWD/gpurun_out/r2_test_values.jsonl
rm -f $DEPGAN_TEST_LOG
timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/r2_gpu_tests_final.log 2>&1
echo "gpu tests exit $?"; tail -n 6 gpurun_out/r2_gpu_tests_final.log
unset DEPGAN_TEST_LOG
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err
echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
echo "ref exit $?"; tail -c 600 gpurun_out/bench_r2_final_ref.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final_1gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], d["clocks"])
print("train", d["train"]["value"], d["train"]["ms_per_step"])
for k in ("configs[0]", "configs[3]", "configs[4]"):
    v = d["configs"][k]
    print(k, {kk: v[kk] for kk in v if kk in ("value", "ms_per_step", "unit")} if isinstance(v, dict) else v)
print("cpu", d["cpu_baseline"])
print("pv", d.get("precision_variants"))
print("predict", d.get("predict_numpy"))
PY
