#!/bin/bash
# Round evidence on the committed code (one GPU call): GPU tests, both bench arms, ncu launch list of bench.py,
# ncu --set full over one whole frame and over the training-iteration kernels.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ev_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/ev_pytest.log
python bench.py --impl reference > gpurun_out/ev_bench_reference.json 2> gpurun_out/ev_bench_reference.err
python bench.py > gpurun_out/ev_bench_ours.json 2> gpurun_out/ev_bench_ours.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ev_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ev_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ev_ncu_bench.log 2>&1
python tools/dbg_step.py > gpurun_out/ev_plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -s 65 -c 13 -o gpurun_out/ev_frame python tools/dbg_step.py > gpurun_out/ev_ncu_frame.log 2>&1
python tools/dbg_train.py > gpurun_out/ev_plain_train.log 2>&1 &&
ncu --set full --clock-control none -k regex:'ssim|adam|densify|l1_only|preprocess_lonlat' -s 20 -c 8 -o gpurun_out/ev_train python tools/dbg_train.py > gpurun_out/ev_ncu_train.log 2>&1
tail -3 gpurun_out/ev_pytest.log; cat gpurun_out/ev_bench_ours.json | cut -c1-400; tail -2 gpurun_out/ev_ncu_frame.log
