#!/bin/bash
# gpurun with retries while the pod answers "transient" (nothing charged): usage gpurun_retry.sh <log> <timeout> <command...>
log="$1"; shift; to="$1"; shift
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|answers busy\|exit code 3" "$log"; then break; fi
  sleep 150
done
