#!/bin/bash
# multi-GPU bench exactly as the driver launches it:  gpurun --gpus N -- 'bash tools/bench_multi_gpu.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/dp_bench_n$N.json 2> gpurun_out/dp_bench_n$N.err
echo "bench exit $?"; tail -5 gpurun_out/dp_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/dp_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'))
print(d['config']['gradient_exchange']); print('exchange_check', d.get('exchange_check'))
for k in ("dp_views","bands"):
    if k in d: print(k, d[k]['value'], d[k]['unit'], d[k].get('ms_per_step'), json.dumps(d[k]['config'])[:400])
PY
if [ "$2" = "dense" ]; then
OGS_DP_EXCHANGE=dense timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 100 --warmup 5 --no-extra > gpurun_out/dp_bench_dense_n$N.json 2> gpurun_out/dp_bench_dense_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/dp_bench_dense_n$N.json').read().strip().splitlines()[-1]); print('dense N', d['n_gpus'], d['value'], d['ms_per_step'])"
fi
