#!/bin/bash
# gpurun with retries while the pod is busy (nothing is charged for a refused call)
#   tools/gpurun_retry.sh [--gpus N] TIMEOUT COMMAND...
gp=()
if [ "$1" == "--gpus" ]; then gp=(--gpus "$2"); shift 2; fi
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "${gp[@]}" --timeout "$t" -- "$@" > /tmp/gpurun_try.log 2>&1
  rc=$?
  if grep -q "status=transient\|retry in a few minutes" /tmp/gpurun_try.log || [ $rc -eq 3 ]; then
    echo "[retry $i] pod busy (rc=$rc), sleeping 90 s"; sleep 90; continue
  fi
  cat /tmp/gpurun_try.log | tail -80
  exit $rc
done
echo "gave up"; exit 3
