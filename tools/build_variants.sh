#!/bin/bash
# builds tuning variants of the library: tools/build_variants.sh name1 "flags1" name2 "flags2" ...
set -e
cd "$(dirname "$0")/../snake-vk_b200"
while [ $# -gt 1 ]; do
  make -s -j8 VARIANT=_$1 EXTRA="$2" && echo "built libsnkpt_$1.so ($2)"
  shift 2
done
