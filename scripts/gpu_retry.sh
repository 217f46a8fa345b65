#!/bin/bash
# usage: gpu_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod answers "busy")
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  echo "$out" | tail -60
  if ! echo "$out" | grep -q "status=transient\|exit code 3\|rc=3"; then break; fi
  echo "[gpu_retry] busy, retry $i"; sleep 90
done
