#!/bin/bash
# gpurun --gpus N with retries while the pod is busy: usage gpurun_retry_n.sh <N> <log> <timeout> <command...>
n="$1"; shift; log="$1"; shift; to="$1"; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --gpus "$n" --timeout "$to" -- "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|status=busy\|rc=3\b" "$log" && ! grep -q "no box\|busy" "$log"; then break; fi
  sleep 180
done
