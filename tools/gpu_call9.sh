#!/bin/bash
mkdir -p gpurun_out
{
for c in C2 C4 C5; do
echo "== hit bytes $c";   python tools/stage_times.py $c
echo "== own tests $c";  OGS_BWD_HITS=0 python tools/stage_times.py $c
done
} > gpurun_out/c9_variants.log 2>&1
cat gpurun_out/c9_variants.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c9_pytest.log; tail -5 gpurun_out/c9_pytest.log
