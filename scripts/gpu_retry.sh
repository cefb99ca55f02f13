#!/bin/bash
# local helper (runs in the build container, not on the GPU box): retries `gpurun` while the pod answers "busy / draining"
# usage: scripts/gpu_retry.sh <timeout-seconds> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then
    echo "[gpu_retry] attempt $i: pod busy, retrying in 45 s" >&2
    sleep 45
    continue
  fi
  echo "$out"
  exit 0
done
echo "[gpu_retry] gave up after 40 attempts"
exit 3
