#!/bin/bash
# one gpurun call: ncu --set full of selected kernels of the north-star frame; usage: KERNEL=regex SKIP=n COUNT=n TAG=x bash tools/run_ncu.sh
TAG=${TAG:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_trace_q} -s ${SKIP:-14} -c ${COUNT:-4} -f -o $OUT/prof \
  env MODE=northstar FRAMES=2 SNKPT_NS_OVERLAP=0 python tools/quick_bench.py ${SCENES:-c2} > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $OUT/ncu.log
ls -la $OUT
