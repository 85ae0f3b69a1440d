#!/bin/bash
# usage: tools/gpu_retry.sh LOGFILE TIMEOUT CMD...   (retries while the pod answers busy/transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient\|rc=3\|busy" $log && ! grep -q "status=ok\|status=done\|exit code" $log; then sleep 60; continue; fi
  if [ $rc -eq 3 ]; then sleep 60; continue; fi
  break
done
echo "gpu_retry finished rc=$rc" >> $log
