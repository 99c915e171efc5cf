#!/bin/bash
# usage: gpurun_retry.sh <outfile> <gpurun args...>   — retries while the pod reports a transient busy state
out="$1"; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if grep -q "status=transient" "$out" || grep -q "retry in a few minutes" "$out"; then sleep 150; else break; fi
done
