#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (small shapes; slow under the tool).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  echo "=== $*"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest "$@" -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -12
}
run tests/test_gpu_ops.py -k "fps or farthest or sample"
run tests/test_gpu_scorenet.py -k "fused_operand and 6144"
run tests/test_gpu_conv_train.py -k "level0 or linear_first"
