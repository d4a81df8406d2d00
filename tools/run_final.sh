#!/bin/bash
# final evidence of a build: GPU tests, smoke, bench (+ reference arm), ncu launch list, ncu --set full digest of one north-star frame
TAG=${1:-final}
bash tools/gpu_check.sh $TAG tests smoke bench benchref launches 2>&1 | tail -12
mkdir -p /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_primary|k_rs_|k_ns_|k_trace_q|k_build_small|k_instance_bounds' \
  -s 66 -c 24 -f -o /tmp/ncu/prof_frame env MODE=northstar FRAMES=1 SNKPT_NS_OVERLAP=0 python tools/quick_bench.py c2 > gpurun_out/$TAG/ncu_frame.log 2>&1; echo "ncu frame rc=$?"
python tools/ncu_digest.py /tmp/ncu/prof_frame.ncu-rep gpurun_out/$TAG/prof_frame.txt 40
for k in k_trace_q k_ns_vertex k_rs_spatial k_rs_temporal k_rs_final; do
  cuobjdump -sass -fun $(cuobjdump -elf snake-vk_b200/libsnkpt.so 2>/dev/null | grep -o "_ZN3snk[A-Za-z0-9_]*${k}[EI][A-Za-z0-9_]*" | sort -u | head -1) snake-vk_b200/libsnkpt.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > gpurun_out/$TAG/$k.sass
done
ls gpurun_out/$TAG
