#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_conv_rowg.py tests/test_gpu_conv.py -m gpu -x -q --timeout 600 > gpurun_out/r2_tests13.log 2>&1
echo "tests exit $?"; tail -n 6 gpurun_out/r2_tests13.log
C="tc_3x3_32to32_plain_N64 tc_3x3_32to32_pool_N64 tc_3x3_64to64_pool_N64"
echo "== default (conv_row first)"; timeout 200 python scripts/kbench.py $C
echo "== DEPGAN_NO_ROW=1 (rowg takes 32->32)"; DEPGAN_NO_ROW=1 timeout 200 python scripts/kbench.py $C
echo "== DEPGAN_NO_ROW=1 DEPGAN_NO_ROWG=1 (tile kernel)"; DEPGAN_NO_ROW=1 DEPGAN_NO_ROWG=1 timeout 200 python scripts/kbench.py $C
timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_pf.json 2> gpurun_out/bench_r2_pf.err
DEPGAN_NO_INFER_POOLFUSE=1 timeout 600 python bench.py --no-train --no-extra --no-cpu > gpurun_out/bench_r2_nopf.json 2> gpurun_out/bench_r2_nopf.err
python - <<'PY'
import json
for f in ("bench_r2_pf", "bench_r2_nopf"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["dtype"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_classes_ms_per_step"])
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
PY
