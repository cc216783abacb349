#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod reports "busy" (exit 3)
log=$1; shift
for i in $(seq 1 20); do
  gpurun "$@" > "$log" 2>&1
  rc=$?
  echo "exit=$rc attempt=$i" >> "$log"
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
