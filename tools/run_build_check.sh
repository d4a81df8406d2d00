#!/bin/bash
# sort + refit changes: correctness first (every step under its own timeout), then the C5 sweep, then a launch list of one 10 M build
OUT=gpurun_out/r07c; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sort.py -m gpu -x -q 2>&1 | tail -15
timeout 400 python -m pytest tests/test_gpu_trace.py tests/test_gpu_particles.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/build_sweep.py 10000000 | grep blas | tee $OUT/build_sweep.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_10M.csv env ONLY=1e7 python tools/build_sweep.py 10000000 > /dev/null 2>&1
python tools/launch_means.py $OUT/launches_10M.csv 2>&1 | tail -30
