#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "transient" (nothing charged)
log=$1; shift; to=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 90
done
exit 3
