#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r2_final_4gpu.json 2> gpurun_out/bench_r2_final_4gpu.err
echo "bench exit $?"; tail -n 3 gpurun_out/bench_r2_final_4gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2_final_4gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e_float16_output"]["value"]), "frac", d["roofline"]["frac"])
print("train", d["train"]["value"], "c3", d["configs"]["configs[3]"]["value"], "c4", d["configs"]["configs[4]"]["value"], d["configs"]["configs[4]"]["labels_only"]["slice_forwards_per_s"])
PY
