#!/bin/bash
OUT=gpurun_out/r07d; mkdir -p $OUT
timeout 200 python tools/timeline.py c2 2>&1 | tee $OUT/timeline_default.txt
for v in "-1,1" "-1,2" "-1,3" "1,2" "2,2" "2,3" "0,2"; do
  SNKPT_NS_STAGGER=$v TIMELINE=${TL:-0} timeout 200 python tools/timeline.py c2 2>&1 | tee -a $OUT/stagger_variants.txt
done
SNKPT_NS_STAGGER=-1,2 timeout 200 python tools/timeline.py c2 2>&1 > $OUT/timeline_stagger_m1_2.txt
