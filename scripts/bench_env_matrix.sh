#!/bin/bash
# Pipelined step time under combinations of the plan's environment switches (one line per combination).
cd "$(dirname "$0")/.."
run() {
  echo -n "$* : "
  env "$@" timeout 300 python scripts/fused_a_latency.py 2>&1 | grep "mode 2" | cut -c1-100
}
run A=0
run REGNET_FPS_CORUN_SINGLE=0
run REGNET_DEFER_PREFETCH=1 REGNET_FPS_CORUN_SINGLE=0
run REGNET_DEFER_PREFETCH=2 REGNET_FPS_CORUN_SINGLE=0
run REGNET_DEFER_PREFETCH=2 REGNET_FPS_CORUN_SINGLE=0 REGNET_SA_FUSED_A=1
run REGNET_DEFER_PREFETCH=1 REGNET_FPS_CORUN_SINGLE=0 REGNET_SA_FUSED_A=1
