#!/bin/bash
# Build kernel-experiment variants of libptg_b200.so into variants/ (git-ignored, travels with gpurun):
#   tools/build_variants.sh name1:"-DFOO=1 -DBAR=0" name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
    name="${spec%%:*}"; flags="${spec#*:}"
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --shared -Xcompiler -fPIC \
         -Xcompiler -O2 $flags -o variants/$name.so rl_ptg_b200/csrc/ptg_capi.cu -ldl &
done
wait
ls -la variants/*.so
