#!/bin/bash
# submit.sh <log> <timeout_s> [--gpus N] -- <command>: gpurun with retries while the pod answers "busy" (exit 3)
LOG=$1; shift; TMO=$1; shift
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift; shift; fi
[ "$1" == "--" ] && shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO $EXTRA -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
