#!/bin/bash
# Retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  Usage:
#   tools/gpurun_retry.sh <timeout_s> <logfile> '<command>' [extra gpurun flags...]
T=$1; LOG=$2; CMD=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] attempt $i finished rc=$rc" >> "$LOG"; exit $rc; fi
  sleep 45
done
echo "[retry] gave up after 40 busy answers" >> "$LOG"; exit 3
