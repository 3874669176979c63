#!/bin/bash
# ncu --set full of the training kernels added late in round 2 (level-0 recompute, linear-first gather / scatter).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k "regex:sa0_input|sa0_apply|sa0_backward_sums|gather_linear|scatter_linear|fp_dense_wgrad" -s 8 -c 8 -o gpurun_out/r02_prof_train_new python scripts/one_train_step.py > gpurun_out/r02_prof_train_new.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r02_prof_train_new.ncu-rep > gpurun_out/r02_prof_train_new_full.txt
rm -f gpurun_out/r02_prof_train_new.ncu-rep
cut -c1-330 gpurun_out/r02_prof_train_new_full.txt
