#!/bin/bash
# r02 evidence on the committed code (one 1-GPU call): both bench arms, ncu launch list of bench.py, ncu --set full over one C2 frame.
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference > gpurun_out/r02ev_bench_reference.json 2> gpurun_out/r02ev_bench_reference.err
timeout 900 python bench.py > gpurun_out/r02ev_bench_ours.json 2> gpurun_out/r02ev_bench_ours.err; echo "bench exit $?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02ev_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02ev_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02ev_ncu_bench.log 2>&1
python tools/dbg_step.py > gpurun_out/r02ev_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 75 -c 15 -o gpurun_out/r02ev_frame python tools/dbg_step.py > gpurun_out/r02ev_ncu_frame.log 2>&1
cut -c1-300 gpurun_out/r02ev_bench_ours.json; tail -2 gpurun_out/r02ev_ncu_frame.log
