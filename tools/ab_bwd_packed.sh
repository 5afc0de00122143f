#!/bin/bash
# A/B run (one GPU call): packed-FFMA2 render_bwd variants, tile_ranges on the side stream, FFMA2 yardstick.
mkdir -p gpurun_out
{
echo "== ffma2 yardstick"; tools/yardstick/ffma2_bench
echo "== gradient errors, packed (default)"; python tools/grad_err.py
echo "== gradient errors, scalar"; OGS_BWD_PACKED=0 python tools/grad_err.py
for v in "OGS_BWD_PACKED=0 OGS_BWD_MINBLOCKS=16 OGS_SIDE_STREAM=0" "OGS_BWD_PACKED=0 OGS_BWD_MINBLOCKS=14 OGS_SIDE_STREAM=0" \
         "OGS_BWD_PACKED=1 OGS_BWD_MINBLOCKS=16 OGS_SIDE_STREAM=0" "OGS_BWD_PACKED=1 OGS_BWD_MINBLOCKS=14 OGS_SIDE_STREAM=0" \
         "OGS_BWD_PACKED=1 OGS_BWD_MINBLOCKS=12 OGS_SIDE_STREAM=0" "OGS_BWD_PACKED=1 OGS_BWD_MINBLOCKS=10 OGS_SIDE_STREAM=0" \
         "OGS_BWD_PACKED=1 OGS_BWD_MINBLOCKS=12 OGS_SIDE_STREAM=1" "OGS_BWD_PACKED=0 OGS_BWD_MINBLOCKS=16 OGS_SIDE_STREAM=1"; do
  echo "== $v"; env $v python tools/stage_times.py C2
done
} > gpurun_out/ab_bwd_packed.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/ab_pytest.log
tail -5 gpurun_out/ab_pytest.log; cat gpurun_out/ab_bwd_packed.log
