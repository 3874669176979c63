#!/bin/bash
# Round-2 ncu captures of one lone forward (fused SA operands) and of the materialised form the pipelined step runs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_train.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/train_step.py --steps 10 2>/dev/null | cut -c100-250
timeout 900 ncu --set full --clock-control none -k "regex:gemm_tc|sa0_chain|gemm_fused_a" -s 20 -c 20 -o gpurun_out/r02_prof_tensor python scripts/one_forward.py tc serial > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
REGNET_SA_FUSED_A=3 timeout 900 ncu --set full --clock-control none -k "regex:gemm_tc|sa0_chain|gemm_fused_a" -s 20 -c 20 -o gpurun_out/r02_prof_tensor_materialised python scripts/one_forward.py tc serial > gpurun_out/r02_ncu_full_m.log 2>&1; tail -2 gpurun_out/r02_ncu_full_m.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_forward.csv python scripts/one_forward.py tc serial > /dev/null 2>&1
timeout 300 python scripts/e2e_profile.py > gpurun_out/r02_e2e_profile.txt 2>&1; tail -5 gpurun_out/r02_e2e_profile.txt | cut -c1-200
ls -la gpurun_out/*.ncu-rep
