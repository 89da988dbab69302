#!/bin/bash
# usage: [GPUS=N] scripts/gpu_retry.sh <timeout_s> <logfile> <command>   -- retries while the pod answers busy (nothing is charged then)
T=$1; LOG=$2; shift 2
cd /root/repo
G=""
[ -n "$GPUS" ] && G="--gpus $GPUS"
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -60 $LOG
