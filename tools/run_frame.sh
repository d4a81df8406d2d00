#!/bin/bash
# one gpurun call: GPU tests, then the per-kernel table (profiling = single stream) and the un-profiled frame time with / without chain overlap
TAG=${TAG:-fr}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q ${TESTS:-} > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
fi
for ov in 1 0; do
  echo "#### SNKPT_NS_OVERLAP=$ov"
  SNKPT_NS_OVERLAP=$ov MODE=northstar FRAMES=${FRAMES:-4} timeout 300 python tools/quick_bench.py ${SCENES:-c2} > $OUT/quick_ov$ov.log 2>&1; cat $OUT/quick_ov$ov.log
done
