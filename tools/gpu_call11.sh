#!/bin/bash
mkdir -p gpurun_out
V=$PWD/gpurun_variants
{
for c in C2 C4 C5; do
echo "== chunk 256 $c";  python tools/stage_times.py $c
echo "== chunk 512 $c";  OMNIGS_B200_LIB=$V/libomnigs_b200_chunk512.so python tools/stage_times.py $c
echo "== chunk 128 $c";  OMNIGS_B200_LIB=$V/libomnigs_b200_chunk128.so python tools/stage_times.py $c
done
} > gpurun_out/c11_variants.log 2>&1
cat gpurun_out/c11_variants.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c11_pytest.log; tail -5 gpurun_out/c11_pytest.log
