#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> '<command>'  -- gpurun, retried while the pod answers "busy" (nothing charged)
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpu_retry: still busy after 30 attempts"; exit 3
