#!/bin/bash
# Tuning variants that differ in ONE translation unit: tools/build_kernel_variants.sh <unit> name1 "flags1" name2 "flags2" ...
# Reuses the product build's objects (snake-vk_b200/build/*.o) and recompiles only csrc/<unit>.cu with the extra flags.
set -e
cd "$(dirname "$0")/../snake-vk_b200"
unit=$1; shift
make -s -j16 libsnkpt.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -gt 1 ]; do
  d=build_$1; mkdir -p $d; cp build/*.o $d/
  $NVCC $ARCH -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $2 -c csrc/$unit.cu -o $d/$unit.o 2> $d/$unit.ptxas.log || (cat $d/$unit.ptxas.log; exit 1)
  $NVCC $ARCH -shared -o libsnkpt_$1.so $d/*.o
  echo "built libsnkpt_$1.so ($2): $(grep -h 'Used' $d/$unit.ptxas.log | tr '\n' ' ')"
  shift 2
done
