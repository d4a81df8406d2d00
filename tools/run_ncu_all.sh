#!/bin/bash
# one gpurun call: ncu --set full of every kernel of ONE north-star frame on C2 and of the 10 M-triangle build / refit; the reports
# are digested on the box (tools/ncu_digest.py) because they exceed what gpurun brings back; SASS of the hot kernels by cuobjdump
TAG=${TAG:-ncuall}
OUT=gpurun_out/$TAG
mkdir -p $OUT /tmp/ncu
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_primary|k_rs_|k_ns_|k_trace_q|k_build_small|k_instance_bounds' \
  -s ${SKIP:-66} -c ${COUNT:-24} -f -o /tmp/ncu/prof_frame env MODE=northstar FRAMES=1 SNKPT_NS_OVERLAP=0 python tools/quick_bench.py c2 > $OUT/ncu_frame.log 2>&1; echo "ncu frame rc=$?"
python tools/ncu_digest.py /tmp/ncu/prof_frame.ncu-rep $OUT/prof_frame.txt 40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_collapse|k_sweep_pass|k_sweep_hist|k_refit_nodes8|k_refit_binary|k_karras|k_morton|k_triangle_bounds' \
  -s 13 -c 16 -f -o /tmp/ncu/prof_build env ONLY=1e7 python tools/build_sweep.py 10000000 > $OUT/ncu_build.log 2>&1; echo "ncu build rc=$?"
python tools/ncu_digest.py /tmp/ncu/prof_build.ncu-rep $OUT/prof_build.txt 20
for k in k_trace_q k_primary_packet k_rs_initial k_rs_spatial k_rs_final k_ns_vertex k_build_small k_collapse k_sweep_pass k_refit_binary_chunked k_refit_nodes8; do
  cuobjdump -sass -fun $(cuobjdump -elf snake-vk_b200/libsnkpt.so 2>/dev/null | grep -o "_ZN3snk[A-Za-z0-9_]*${k}[EI][A-Za-z0-9_]*" | sort -u | head -1) snake-vk_b200/libsnkpt.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > $OUT/$k.sass
  echo "$k: $(wc -l < $OUT/$k.sass) SASS lines"
done
ls -la $OUT; du -sh $OUT
