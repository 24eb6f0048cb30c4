#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy" (exit code 3)
T=$1; shift
G=${GPUS:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
