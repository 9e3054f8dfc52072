#!/bin/bash
# round-2: 2-GPU tests (data parallel: torch / peer / nccl transports, sync-BN fit) + 2-GPU train bench A/B of the transports
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_infer_dp.py -m gpu -x -q --timeout 900 > gpurun_out/r2_dp_tests.log 2>&1
echo "dp tests exit $?"; tail -n 30 gpurun_out/r2_dp_tests.log
for kind in peer nccl torch; do
  DEPGAN_COLLECTIVE=$kind timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload depgan_train --steps 8 --warmup 3 > gpurun_out/bench_r2_train_2gpu_$kind.json 2> gpurun_out/bench_r2_train_2gpu_$kind.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_train_2gpu_$kind.json").read().strip().splitlines()[-1])
    print("$kind", d["value"], d["ms_per_step"], d.get("dp_check"))
except Exception as e:
    print("$kind failed", e); print(open("gpurun_out/bench_r2_train_2gpu_$kind.err").read()[-1500:])
PY
done
