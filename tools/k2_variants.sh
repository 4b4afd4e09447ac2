#!/bin/bash
# Build A/B variants of the K2 kernel (prune_fused2.cu with ablation macros) as separate libraries under cafe_b200/build/variants/.
# usage: tools/k2_variants.sh NAME "-DFLAG ..." [NAME "-DFLAG" ...]
set -e
cd "$(dirname "$0")/../cafe_b200/csrc"
mkdir -p ../build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -c -o ../build/variants/prune_fused2_$name.o prune_fused2.cu
  objs=$(ls ../build/*.o | grep -v prune_fused2.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../build/variants/libcafe_gpu_$name.so $objs ../build/variants/prune_fused2_$name.o -ldl
  echo built $name
done
