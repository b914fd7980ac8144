#!/bin/bash
# usage: tools/gpu_retry.sh <timeout-seconds> [--gpus N] '<command>'  -- gpurun, retried while the pod answers "busy" (nothing charged)
T=$1; shift
OPTS=""
if [ "$1" == "--gpus" ]; then OPTS="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T $OPTS -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpu_retry: still busy after 40 attempts"; exit 3
