#!/bin/bash
# Round-2 ncu captures of one lone forward (fused SA operands) and of the materialised form the pipelined step runs.
# The reports are summarised on the box (scripts/ncu_summary.py) and deleted: gpurun brings back at most 64 MiB.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {   # name, extra env
  env $2 timeout 900 ncu --set full --clock-control none -k "regex:gemm_tc|sa0_chain|gemm_fused_a" -s 20 -c 20 \
      -o gpurun_out/$1 python scripts/one_forward.py tc serial > gpurun_out/$1.log 2>&1
  python scripts/ncu_summary.py full gpurun_out/$1.ncu-rep > gpurun_out/$1_full.txt
  python scripts/ncu_summary.py traffic gpurun_out/$1.ncu-rep > gpurun_out/$1_traffic.json
  rm -f gpurun_out/$1.ncu-rep
  tail -3 gpurun_out/$1_traffic.json
}
cap r02_prof_tensor REGNET_SA_FUSED_A=2
cap r02_prof_tensor_materialised REGNET_SA_FUSED_A=3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_forward.csv python scripts/one_forward.py tc serial > /dev/null 2>&1
timeout 300 python scripts/e2e_profile.py > gpurun_out/r02_e2e_profile.txt 2>&1; tail -5 gpurun_out/r02_e2e_profile.txt | cut -c1-200
du -sh gpurun_out
