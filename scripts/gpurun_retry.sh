#!/usr/bin/env bash
# usage: gpurun_retry.sh <timeout-seconds> '<command>'   - retries while the pod answers "busy" (exit code 3)
t="$1"; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
