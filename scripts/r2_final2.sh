#!/bin/bash
# round-2 final check on two GPUs: the default bench line under torchrun (inference + train legs + configs[3] strong scaling +
# cohort sweep), as the driver launches it
mkdir -p gpurun_out
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_2gpu.json 2> gpurun_out/bench_r2_final_2gpu.err
echo "bench exit $?"; tail -n 5 gpurun_out/bench_r2_final_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final_2gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], d["config"].get("numa_binding_rank0"))
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"].get("dp_check"))
for k in ("configs[0]", "configs[3]", "configs[4]"):
    v = d["configs"][k]
    print(k, {kk: v[kk] for kk in v if kk in ("value", "ms_per_step", "unit", "dp_check")} if isinstance(v, dict) else v)
print("predict", d.get("predict_numpy"))
PY
timeout 600 python -m pytest tests/test_gpu_infer_dp.py -m gpu -x -q --timeout 900 2>&1 | tail -3
