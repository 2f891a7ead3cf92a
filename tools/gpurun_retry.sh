#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...>  -- retries while the pod answers busy / transient (nothing charged)
g=$1; t=$2; shift 2
for i in $(seq 1 20); do
  out=$(gpurun --gpus $g --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|exit code 3\|status=busy\|no box"; then sleep 120; continue; fi
  break
done
