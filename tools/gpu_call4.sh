#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/c4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c4_pytest.log
tail -22 gpurun_out/c4_pytest.log
V=$PWD/gpurun_variants
{
for c in C2 C3 C4 C5; do
echo "== segments $c";  python tools/stage_times.py $c
echo "== classic $c";   OGS_SEGMENT_SORT=0 python tools/stage_times.py $c
done
} > gpurun_out/c4_variants.log 2>&1
cat gpurun_out/c4_variants.log
OMNIGS_B200_LIB=$V/libomnigs_b200_timeline.so timeout 600 python tools/tile_timeline.py C2 C4 C5 > gpurun_out/r02_tile_timeline.json 2> gpurun_out/c4_timeline.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_tile_timeline.json'))
for c,v in d.items():
    for o,r in v.items():
        if 'error' in r: print(c,o,r); continue
        for k in ('render_fwd','render_bwd'):
            x=r[k]; print(c,o,k,{a:(round(b,3) if isinstance(b,float) else b) for a,b in x.items() if a!='row_mean_tile_us'}, x.get('row_mean_tile_us'))
PY
timeout 900 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench exit $?"; tail -2 gpurun_out/c4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c4_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
print(d['roofline_frame']['stage_ms'])
print('dp_views', d['dp_views']['ms_per_step'], 'bands', d['bands']['value'])
PY
