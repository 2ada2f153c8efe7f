#!/bin/bash
# Builds experiment variants of the library into variants/ (git-ignored, travels with gpurun):
#   scripts/build_variants.sh name "-DWS_FLOOD_STAGES=3" [name2 "flags2" ...]
# Run one with WS_B200_LIB=variants/libws_<name>.so.
cd "$(dirname "$0")/.." || exit 1
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
    -o variants/libws_$name.so rustronomy-watershed_b200/csrc/{kernels,flood,labels,merge,forest,engine}.cu || exit 1
  echo "built variants/libws_$name.so ($flags)"
done
