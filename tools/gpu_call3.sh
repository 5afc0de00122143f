#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/c3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c3_pytest.log
tail -22 gpurun_out/c3_pytest.log
V=$PWD/gpurun_variants
{
echo "== default";            python tools/stage_times.py C2
echo "== staging ldgsts";     OGS_FWD_STAGING=ldgsts python tools/stage_times.py C2
echo "== staging bulk";       OGS_FWD_STAGING=bulk python tools/stage_times.py C2
echo "== staging ldgsts, 8 CTAs/SM"; OMNIGS_B200_LIB=$V/libomnigs_b200_async8.so OGS_FWD_STAGING=ldgsts python tools/stage_times.py C2
echo "== staging bulk, 8 CTAs/SM";   OMNIGS_B200_LIB=$V/libomnigs_b200_async8.so OGS_FWD_STAGING=bulk python tools/stage_times.py C2
echo "== preprocess_bwd 6 CTAs/SM";  OMNIGS_B200_LIB=$V/libomnigs_b200_pb6.so python tools/stage_times.py C2
echo "== default C4";         python tools/stage_times.py C4
echo "== staging ldgsts C4";  OGS_FWD_STAGING=ldgsts python tools/stage_times.py C4
echo "== default C5";         python tools/stage_times.py C5
} > gpurun_out/c3_variants.log 2>&1
cat gpurun_out/c3_variants.log
echo "== parity of the async staging kernels"
OGS_FWD_STAGING=ldgsts timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "golden or against_reference_rasterizer or c2 or tiny or seam" 2>&1 | tail -3
OGS_FWD_STAGING=bulk timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_seam_wrap.py -m gpu -x -q -k "golden or against_reference_rasterizer or c2 or tiny or seam" 2>&1 | tail -3
