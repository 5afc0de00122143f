#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{
echo "== bands cost-balanced, halo, 4 ranges"; timeout 300 $TR --master-port 29541 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-1100
echo "== bands cost-balanced, halo, 1 range";  OGS_BAND_CHUNKS=1 timeout 300 $TR --master-port 29542 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-1100
echo "== bands instance-balanced, halo, 4 ranges"; OGS_BAND_BALANCE=instances timeout 300 $TR --master-port 29543 tools/bench_blocks.py bands 2>&1 | grep '^{' | cut -c1-1100
} > gpurun_out/bands_variants2_n$N.log 2>&1
cat gpurun_out/bands_variants2_n$N.log
{
for cap in 0 296 148 74 37; do
for defer in 1 0; do
echo "== rebuild blocks cap $cap defer $defer"
OGS_SH_REBUILD_BLOCKS=$cap OGS_DP_DEFER=$defer timeout 300 $TR --master-port 29545 tools/dp_trace.py 40 2>/dev/null | grep '^{' | head -1 | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print(round(d['ms_per_step'],4), {k:round(v) for k,v in d['gpu_us'].items()}, d['stage_us'])"
done; done
} > gpurun_out/dp_rebuild_sweep_n$N.log 2>&1
cat gpurun_out/dp_rebuild_sweep_n$N.log
