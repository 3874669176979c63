#!/bin/bash
# Everything one gpurun call should do; each stage under its own timeout, logs into gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
python -c "import os; print('cores', os.cpu_count())" >> gpurun_out/smi.txt
STAGES="${1:-golden ops mlp scorenet smoke bench sweep ncu}"
for s in $STAGES; do
  echo "=== stage $s $(date +%T)"
  case $s in
    golden)   timeout 600 python oracle/gen_golden_gpu.py gpurun_out/golden > gpurun_out/golden.log 2>&1 ;;
    ops)      timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/test_ops.log 2>&1 ;;
    mlp)      timeout 600 python -m pytest tests/test_gpu_mlp.py -q -m gpu -s > gpurun_out/test_mlp.log 2>&1 ;;
    scorenet) timeout 900 python -m pytest tests/test_gpu_scorenet.py -q -m gpu > gpurun_out/test_scorenet.log 2>&1 ;;
    smoke)    timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ;;
    bench)    timeout 900 python bench.py > gpurun_out/bench_tc.log 2>&1 ;;
    benchsimt) timeout 600 python bench.py --steps 5 --warmup 3 --engine simt --no-cpu-baseline > gpurun_out/bench_simt.log 2>&1 ;;
    refcuda)  timeout 900 python oracle/bench_ref_cuda.py > gpurun_out/ref_cuda_baseline.json 2> gpurun_out/ref_cuda_baseline.log ;;
    benchref) timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1 ;;
    sweep)    timeout 300 python scripts/fps_sweep.py > gpurun_out/fps_sweep.log 2>&1 ;;
    timeline) timeout 300 python scripts/timeline.py 2 > gpurun_out/timeline_mode2.log 2>&1
              
              ;;
    corun)    for v in 8,128 8,256 8,512; do REGNET_FPS_CORUN=$v timeout 300 python scripts/timeline.py 2 > gpurun_out/timeline_corun_${v/,/_}.log 2>&1; done ;;
    ab)       timeout 300 python scripts/pipeline_ab.py > gpurun_out/pipeline_ab.log 2>&1 ;;
    region)   timeout 600 python -m pytest tests/test_gpu_region.py -q -m gpu > gpurun_out/test_region.log 2>&1 ;;
    ncu)      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
                 --log-file gpurun_out/launches.csv python scripts/one_forward.py tc serial > gpurun_out/ncu_list.log 2>&1 ;;
    ncufull)  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc|sa0_chain" -s 20 -c 20 \
                 -o gpurun_out/prof_tensor python scripts/one_forward.py tc serial > gpurun_out/ncu_full.log 2>&1
              timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_kernel -s 3 -c 1 \
                 -o gpurun_out/prof_fps python scripts/one_forward.py tc serial > gpurun_out/ncu_full_fps.log 2>&1
              timeout 600 ncu --set full --clock-control none --import-source on -k regex:affine_warp -c 3 \
                 -o gpurun_out/prof_affine python scripts/one_forward.py tc serial > gpurun_out/ncu_full_affine.log 2>&1 ;;
    sa0chk)   timeout 600 python scripts/sa0_chain_check.py > gpurun_out/sa0_chain_check.log 2>&1 ;;
    bench3)   REGNET_FUSE_SA0=3 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chain.log 2>&1 ;;
    side2)    timeout 300 python scripts/timeline.py 2 > gpurun_out/timeline_side2.log 2>&1
              timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_side2.log 2>&1
              REGNET_NO_SIDE2=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_noside2.log 2>&1
              for v in 1,256 2,256 4,128 4,256; do REGNET_FPS_CORUN1=$v timeout 300 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/bench_side2_c1_${v/,/_}.log 2>&1; done ;;
    opsgrid)  timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/test_ops.log 2>&1 ;;
    e2e)      timeout 300 python scripts/e2e_inference.py > gpurun_out/e2e_c3.json 2> gpurun_out/e2e_c3.err ;;
    alltests) timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/test_all.log 2>&1 ;;
  esac
  echo "    exit $?"
done
tail -n 5 gpurun_out/*.log
