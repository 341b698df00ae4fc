#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3: nothing charged); usage: tools/gpurun_retry.sh TIMEOUT 'command'
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > gpurun_out/last_gpurun.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/last_gpurun.log; then break; fi
  sleep 60
done
tail -30 gpurun_out/last_gpurun.log
