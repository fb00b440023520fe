#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpus> <timeout_s> <command...>   -- retries while the pod answers "busy" (nothing is charged then)
G=$1; T=$2; shift 2
python -m ramscb_b200.build > /dev/null 2>&1 || { echo "BUILD FAILED"; exit 1; }
for try in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "gave up after 40 tries"; echo "$out" | tail -5
