#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy / transient" (nothing is charged for those):
#   [GPURUN_OPTS="--gpus 8"] scripts/gpurun_retry.sh <logfile> <timeout_s> '<command>'
log=$1; to=$2; shift 2
for i in $(seq 1 ${GPURUN_TRIES:-40}); do
  gpurun --timeout "$to" ${GPURUN_OPTS:-} -- "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|retry in a few minutes\|retry later" "$log"; then sleep 45; continue; fi
  break
done
echo "[retry] finished after $i attempt(s)" >> "$log"
