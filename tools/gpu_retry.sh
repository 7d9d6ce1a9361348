#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <command...>   -- retries while the pod answers busy (exit 3 / transient)
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] busy, sleeping 120s"; sleep 120; continue; fi
  exit $rc
done
exit 3
