#!/bin/bash
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in side main side main; do
OGS_DP_COLOURS=$mode timeout 600 $TR --master-port 29551 bench.py --gpus $N --steps 100 --warmup 5 --no-extra 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d.get('exchange_check',{}).get('replicas_bitwise_equal'))"
done
