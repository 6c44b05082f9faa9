#!/usr/bin/env bash
# Retry a gpurun call while the pod answers "busy" (exit code 3: nothing charged).  usage: tools/gpu_retry.sh <timeout_s> <log> <command...>
T=$1; LOG=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$LOG" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
