#!/bin/bash
# round 2, GPU call 1: full GPU test suite, gradient-noise sweep (default + exact-math build), quick bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -15 gpurun_out/c1_pytest.log
timeout 900 python tools/grad_noise.py --tag default --out gpurun_out/r02_grad_noise.json > gpurun_out/c1_noise_default.log 2>&1; echo "exit $?" >> gpurun_out/c1_noise_default.log
OMNIGS_B200_LIB=$PWD/gpurun_variants/libomnigs_b200_exact.so timeout 900 python tools/grad_noise.py --tag exact_math --out gpurun_out/r02_grad_noise.json > gpurun_out/c1_noise_exact.log 2>&1; echo "exit $?" >> gpurun_out/c1_noise_exact.log
tail -60 gpurun_out/c1_noise_default.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench exit $?"
cut -c1-600 gpurun_out/c1_bench.json
