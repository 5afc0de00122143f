#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_spread.py gpurun_out/r02_parity_spread.json 6 > gpurun_out/c6_spread.log 2>&1; tail -80 gpurun_out/c6_spread.log
{
for c in C2 C4 C5; do
echo "== default $c";  python tools/stage_times.py $c
echo "== fwd list $c";   OGS_FWD_LIST=1 python tools/stage_times.py $c
done
} > gpurun_out/c6_variants.log 2>&1
cat gpurun_out/c6_variants.log
echo "== parity with the list kernel"
OGS_FWD_LIST=1 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_seam_wrap.py -m gpu -q -k "golden or against_reference_rasterizer or c2 or tiny or seam or oracle" 2>&1 | tail -8
