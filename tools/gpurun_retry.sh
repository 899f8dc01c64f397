#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>' -- retries while the pod answers busy/draining (rc 3 / transient)
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > gpurun_out/.retry_last.txt 2>&1; rc=$?
  if grep -q "status=ok\|status=fail\|status=timeout" gpurun_out/.retry_last.txt; then cat gpurun_out/.retry_last.txt; exit $rc; fi
  tail -2 gpurun_out/.retry_last.txt; sleep 120
done
exit 3
