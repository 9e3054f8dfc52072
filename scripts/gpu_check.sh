#!/bin/bash
# Runs each GPU test file in its own process (a trapped kernel kills only that file's context) and keeps logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for f in "$@"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -m gpu -x -q --timeout 300 > "gpurun_out/${name}.log" 2>&1
  r=$?
  echo "== $f -> exit $r"; tail -n 25 "gpurun_out/${name}.log"
  [ $r -ne 0 ] && rc=$r
done
exit $rc
