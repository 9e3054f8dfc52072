#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_8gpu.json 2> gpurun_out/bench_r2_final_8gpu.err
echo "bench exit $?"; tail -n 3 gpurun_out/bench_r2_final_8gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final_8gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], d["config"].get("numa_binding_rank0"))
print("train", d["train"]["value"], d["train"]["ms_per_step"], d["train"].get("dp_check"))
for k in ("configs[0]", "configs[3]", "configs[4]"):
    v = d["configs"][k]
    print(k, {kk: v[kk] for kk in v if kk in ("value", "ms_per_step", "unit")} if isinstance(v, dict) else v)
print("predict", d.get("predict_numpy"))
PY
