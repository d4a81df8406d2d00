#!/bin/bash
# one gpurun call: north-star parity tests on the product library, then quick_bench (north-star mode) for the product library with both
# traversal kernels and for every libsnkpt_*.so variant; optional ncu capture of the traversal kernel
TAG=${TAG:-mx}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q ${TESTS:-} > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
fi
export MODE=northstar FRAMES=${FRAMES:-4}
{
  echo "#### product, SNKPT_TRACE_KERNEL=q"
  SNKPT_TRACE_KERNEL=q timeout 300 python tools/quick_bench.py ${SCENES:-c2 c2m} 2>&1
  for lib in "" $(ls snake-vk_b200/libsnkpt_*.so 2>/dev/null | sed 's/.*libsnkpt\(_[^.]*\)\.so/\1/'); do
    echo "#### variant '${lib}'"
    SNKPT_LIB_VARIANT=$lib timeout 300 python tools/quick_bench.py ${SCENES:-c2 c2m} 2>&1
  done
} > $OUT/variants.log 2>&1
grep -E "####|== |frame|trace.q" $OUT/variants.log
if [ -n "$NCU" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_mx -s ${NCU_SKIP:-10} -c ${NCU_COUNT:-3} -f -o $OUT/prof_mx \
    env FRAMES=2 python tools/quick_bench.py ${NCU_SCENE:-c2} > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
fi
ls -la $OUT
