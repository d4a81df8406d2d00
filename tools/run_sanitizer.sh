#!/bin/bash
# one gpurun call: compute-sanitizer memcheck / racecheck / synccheck over the GPU parity tests that exercise every kernel family
# (packet + one-ray traversal, both frame modes, groups, scene front-end, streams); usage: TAG=x bash tools/run_sanitizer.sh
OUT=gpurun_out/${TAG:-sanitizer}; mkdir -p $OUT
T="tests/test_gpu_trace.py tests/test_gpu_frame.py tests/test_gpu_northstar.py tests/test_gpu_group.py tests/test_gpu_scene.py tests/test_gpu_streams.py tests/test_gpu_particles.py tests/test_gpu_taa.py tests/test_gpu_sort.py"
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  t0=$(date +%s)
  timeout ${LIMIT:-700} compute-sanitizer --tool $tool --print-limit 20 python -m pytest $T -q -x -k "not 3000001" > $OUT/$tool.log 2>&1
  echo "$tool rc=$? wall $(( $(date +%s) - t0 )) s: $(grep -E 'passed|failed' $OUT/$tool.log | tail -1) | $(grep -E 'SUMMARY' $OUT/$tool.log | tail -1)"
done | tee $OUT/summary.txt
