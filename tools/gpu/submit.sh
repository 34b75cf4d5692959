#!/bin/bash
# usage: tools/gpu/submit.sh <name> <timeout> <script> [--gpus N]   -- retries while the pod answers "busy" (exit 3, nothing charged)
name=$1; to=$2; script=$3; shift 3
for attempt in $(seq 1 40); do
  gpurun --timeout $to "$@" -- "bash $script" > gpurun_out/${name}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
