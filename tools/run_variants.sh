#!/bin/bash
# runs tools/quick_bench.py for the product library and every libsnkpt_*.so variant (one gpurun call)
mkdir -p gpurun_out/${TAG:-var}
for lib in "" $(ls snake-vk_b200/libsnkpt_*.so 2>/dev/null | sed 's/.*libsnkpt\(_[^.]*\)\.so/\1/'); do
  echo "#### variant '${lib}'"
  SNKPT_LIB_VARIANT=$lib FRAMES=${FRAMES:-4} timeout 300 python tools/quick_bench.py ${SCENES:-c2} 2>&1 
done | tee gpurun_out/${TAG:-var}/variants.log
