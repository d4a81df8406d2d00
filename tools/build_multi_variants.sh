#!/bin/bash
# variants touching several units: name "unit1:flags1;unit2:flags2"
set -e
cd /root/repo/snake-vk_b200
make -s -j16 libsnkpt.so
NVCC=/usr/local/cuda/bin/nvcc; ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -gt 1 ]; do
  d=build_$1; mkdir -p $d; cp build/*.o $d/
  IFS=';' read -ra parts <<< "$2"
  for p in "${parts[@]}"; do
    unit=${p%%:*}; flags=${p#*:}
    fm=""; case $unit in gbuffer|gbuffer_nm|restir|bounce|northstar|restir_ns|taa|particles) fm="-fmad=false";; esac
    $NVCC $ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $fm $flags -c csrc/$unit.cu -o $d/$unit.o 2> $d/$unit.ptxas.log || (cat $d/$unit.ptxas.log; exit 1)
  done
  $NVCC $ARCH -shared -o libsnkpt_$1.so $d/*.o -ldl
  echo "built libsnkpt_$1.so ($2)"
  shift 2
done
