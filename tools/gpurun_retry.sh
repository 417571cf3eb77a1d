#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "transient" / busy
t=$1; shift
for i in $(seq 1 12); do
    out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
    if echo "$out" | grep -q "status=transient\|no box\|busy"; then sleep 150; continue; fi
    echo "$out"; exit 0
done
echo "$out"; exit 3
