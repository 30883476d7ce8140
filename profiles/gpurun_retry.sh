#!/bin/bash
# usage: [GPUS=N] profiles/gpurun_retry.sh <log> <timeout-s> <command...>   - retries while the pod answers "busy" (exit 3)
LOG=$1; shift; TMO=$1; shift
G=""; if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $G --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 60
done
exit 3
