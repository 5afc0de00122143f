#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or oracle or against_reference or multiview or views or raw or train or band or C2 or c2" > gpurun_out/c8_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c8_pytest.log; tail -6 gpurun_out/c8_pytest.log
for c in C2 C3; do python tools/stage_times.py $c; done 2>&1 | tee gpurun_out/c8_stage_times.log
