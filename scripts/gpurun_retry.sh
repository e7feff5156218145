#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout> <command...>  - retries while the pod answers "transient"/busy
T=$1; shift
for i in $(seq 1 30); do
  out=$(gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|no box\|busy"; then
    if echo "$out" | grep -q "status=ok"; then echo "$out"; exit 0; fi
    sleep 90; continue
  fi
  echo "$out"; exit 0
done
echo "gave up after 30 tries"; echo "$out" | tail -5
