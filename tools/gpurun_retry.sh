#!/bin/bash
# usage: [GPUS=N] tools/gpurun_retry.sh <timeout-seconds> '<command>'   (retries while the pod answers busy / no box: rc 3)
T=$1; shift
EXTRA=""
if [ -n "$GPUS" ]; then EXTRA="--gpus $GPUS"; fi
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $EXTRA --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
