#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <log-file> <command string> [--gpus N]
# Retries gpurun while it answers "no box / slot free" (exit 3), every 90 s, for up to ~40 minutes.
to=$1; log=$2; cmd=$3; shift 3
for i in $(seq 1 28); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$to" -- "$cmd" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
