#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
echo "== bands halo + 4 ranges"; timeout 300 $TR --master-port 29541 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-900
echo "== bands halo + 1 range";  OGS_BAND_CHUNKS=1 timeout 300 $TR --master-port 29542 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-900
echo "== bands halo + 8 ranges"; OGS_BAND_CHUNKS=8 timeout 300 $TR --master-port 29543 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-900
echo "== bands full gather + 1 range (first r02 form)"; OGS_BAND_GATHER=full OGS_BAND_CHUNKS=1 timeout 300 $TR --master-port 29544 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-900
} > gpurun_out/bands_variants_n$N.log 2>&1
cat gpurun_out/bands_variants_n$N.log
bash tools/gpu_call_dp.sh $N
timeout 300 $TR --master-port 29545 tools/dp_trace.py 40 2>/dev/null | grep '^{' > gpurun_out/dp_trace_n$N.log; head -c 900 gpurun_out/dp_trace_n$N.log
