#!/bin/bash
# gpurun with retries while the pod answers busy (exit 3): tools/gpurun_retry.sh LOG TIMEOUT 'cmd'
LOG=$1; TMO=$2; CMD=$3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$CMD" > $LOG 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
