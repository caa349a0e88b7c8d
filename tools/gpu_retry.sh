#!/bin/bash
# usage: tools/gpu_retry.sh <tag> <timeout_s> [--gpus N] -- '<command>'   (retries while the pod has no free slot)
tag=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$1" > gpurun_out/${tag}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" gpurun_out/${tag}_call.log; then exit $rc; fi
  sleep 120
done
exit 3
