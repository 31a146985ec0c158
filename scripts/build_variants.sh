#!/bin/bash
# builds alternative versions of the CUDA library for kernel experiments: scripts/build_variants.sh name "-DFLAGS" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p whippet.jl_b200/csrc/variants
while [ $# -gt 1 ]; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $2 \
    -o whippet.jl_b200/csrc/variants/$1.so whippet.jl_b200/csrc/whippet_b200.cu &
  shift 2
done
wait
ls -la whippet.jl_b200/csrc/variants
