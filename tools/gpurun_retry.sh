#!/bin/bash
# usage: [GPURUN_FLAGS='--gpus 2'] tools/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod answers busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 40); do
  out=$(gpurun $GPURUN_FLAGS --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun: still busy after 40 tries"; exit 3
