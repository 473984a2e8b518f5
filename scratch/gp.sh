#!/bin/bash
# retry gpurun while the pod is busy (exit code 3 = nothing charged)
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
