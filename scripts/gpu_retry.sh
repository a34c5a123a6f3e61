#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout> '<command>'   -- retries gpurun while the pod answers busy (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$OUT"; exit 0
done
echo "gpu_retry: gave up"; exit 3
