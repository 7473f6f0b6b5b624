#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout_s> '<command>'   — retries while the pod answers busy (exit 3 / transient), up to 12 times
T=$1; shift
G=""; [ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
