#!/bin/bash
# Pipelined step time under combinations of the plan's environment switches (one line per combination).
cd "$(dirname "$0")/.."
run() {
  echo -n "$* : "
  env "$@" timeout 300 python bench.py --steps 20 --no-train --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'clouds/s', round(d['ms_per_step'],3), 'ms/step  sustained', round(d['sustained']['value'],1), ' latency', round(d['latency_ms_unpipelined'],3))"
}
run A=0
run REGNET_SA_FUSED_A=0
run REGNET_SA_FUSED_A=0 REGNET_DEFER_PREFETCH=1 REGNET_FPS_CORUN_SINGLE=0
run REGNET_SA_FUSED_A=1 REGNET_DEFER_PREFETCH=1 REGNET_FPS_CORUN_SINGLE=0
run REGNET_SA_FUSED_A=1 REGNET_DEFER_PREFETCH=1
run REGNET_SA_FUSED_A=0 REGNET_FPS_CORUN_SINGLE=0
