#!/bin/bash
# quick_bench (north-star mode) for the product library and every libsnkpt_*.so variant; SCENES, TAG
OUT=gpurun_out/${TAG:-var}; mkdir -p $OUT
for lib in "" $(ls snake-vk_b200/libsnkpt_*.so 2>/dev/null | sed 's/.*libsnkpt\(_[^.]*\)\.so/\1/'); do
  echo "#### variant '${lib}'"
  SNKPT_LIB_VARIANT=$lib MODE=northstar FRAMES=${FRAMES:-3} timeout 300 python tools/quick_bench.py ${SCENES:-c2} 2>&1 | grep -E "== |unprofiled|frame |trace.q|primary|build|rs_|ns_"
done > $OUT/variants.log 2>&1
cat $OUT/variants.log
