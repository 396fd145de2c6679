#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <gpus> '<command>'   - retries gpurun while the pod has no free slot
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" -gt 1 ]; then OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); fi
  RC=$?
  if echo "$OUT" | grep -q "status=transient\|no box or slot\|retry in a few minutes"; then sleep 150; continue; fi
  echo "$OUT" | tail -80
  exit $RC
done
echo "gpu_retry: gave up"; exit 3
