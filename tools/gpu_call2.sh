#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/c2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c2_pytest.log
tail -25 gpurun_out/c2_pytest.log
timeout 900 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench exit $?"
tail -3 gpurun_out/c2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c2_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
print(d['roofline_frame']['stage_ms'])
print(json.dumps(d.get('roofline_issue'))[:900])
print(json.dumps(d.get('dp_views'))[:600])
print(json.dumps(d.get('bands'))[:700])
print(d.get('cpu_baseline'))
PY
