#!/bin/bash
# Run each test of a file in its own process under a hard timeout (a hung kernel must not eat the GPU budget).
f=${1:-tests/test_train_step_gpu.py}; lim=${2:-90}
for t in $(python -m pytest "$f" --collect-only -q -m gpu 2>/dev/null | grep "::"); do
  echo "=== $t"
  CUDA_LAUNCH_BLOCKING=1 timeout -k 5 "$lim" python -m pytest "$t" -x -q -m gpu -o faulthandler_timeout=45 2>&1 | tail -25
  echo "rc=$?"
done
