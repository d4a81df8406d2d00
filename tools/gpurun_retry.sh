#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> [--gpus N] -- '<command>'   retries while the pod answers "transient" / busy (rc 3)
log=$1; shift; to=$1; shift
gp=""
if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
shift  # the "--"
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $gp --timeout $to -- "$1" > $log 2>&1
  if grep -q "status=transient\|answers busy\|rc=3" $log || grep -q "no box" $log; then sleep 60; continue; fi
  break
done
