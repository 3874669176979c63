#!/bin/bash
# Round-2 evidence collection on one GPU box (run through gpurun); everything lands in gpurun_out/r02_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests $(date +%T)"; timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_test_all.log 2>&1; tail -3 gpurun_out/r02_test_all.log
echo "=== smoke $(date +%T)"; timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
echo "=== bench $(date +%T)"; timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json
echo "=== bench reference arm $(date +%T)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
echo "=== train profile $(date +%T)"; timeout 300 python scripts/train_profile.py > gpurun_out/r02_train_step_profile.txt 2>/dev/null; head -12 gpurun_out/r02_train_step_profile.txt | cut -c1-160
echo "=== train steps $(date +%T)"; timeout 300 python scripts/train_step.py --steps 10 > gpurun_out/r02_train_step_n1.json 2>/dev/null; cut -c1-260 gpurun_out/r02_train_step_n1.json
REGNET_TRAIN_PASSES=1 timeout 300 python scripts/train_step.py --steps 10 > gpurun_out/r02_train_step_n1_bf16x1.json 2>/dev/null; cut -c100-260 gpurun_out/r02_train_step_n1_bf16x1.json
timeout 300 python scripts/train_full_step.py --steps 10 > gpurun_out/r02_train_full_step_n1.json 2>/dev/null; cut -c1-300 gpurun_out/r02_train_full_step_n1.json
echo "=== C3 $(date +%T)"; timeout 300 python scripts/e2e_inference.py > gpurun_out/r02_e2e_inference_c3.json 2>/dev/null; cat gpurun_out/r02_e2e_inference_c3.json
echo "=== fps $(date +%T)"; timeout 300 python scripts/fps_multi_check.py > gpurun_out/r02_fps_multi_check.txt 2>&1; cat gpurun_out/r02_fps_multi_check.txt
echo "=== ref cuda $(date +%T)"; timeout 900 python oracle/bench_ref_cuda.py > gpurun_out/r02_ref_cuda_baseline.json 2> gpurun_out/r02_ref_cuda_baseline.log; tail -c 700 gpurun_out/r02_ref_cuda_baseline.json
echo "=== reference scripts $(date +%T)"; [ -d baseline/_ref/REGNet ] && timeout 900 python scripts/run_reference_scripts.py > gpurun_out/r02_reference_scripts_on_b200.json 2> gpurun_out/r02_reference_scripts.err
echo "=== done $(date +%T)"
