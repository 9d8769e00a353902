#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient"); usage: tools/gpurun_retry.sh <logfile> <gpurun args...>
log=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
exit $rc
