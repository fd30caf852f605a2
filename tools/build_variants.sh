#!/bin/bash
# usage: bash tools/build_variants.sh name1 "flags1" name2 "flags2" ...   -> ccl_b200/libccl_b200_<name>.so (parallel)
cd "$(dirname "$0")/.."
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
      -I include -o ccl_b200/libccl_b200_$name.so ccl_b200/csrc/*.cu 2> /tmp/build_$name.log \
      && echo "built $name" || { echo "FAILED $name"; tail -5 /tmp/build_$name.log; } ) &
done
wait
