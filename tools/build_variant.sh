#!/bin/bash
# Build a kernel-experiment variant of libptg_b200.so:  tools/build_variant.sh NAME [-DFOO=1 ...]
# The result lands in gpurun_out/../variants/NAME.so (git-ignored, travels with gpurun); select it with PTG_B200_SO.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --shared -Xcompiler -fPIC \
     -Xcompiler -O2 "$@" -o variants/$name.so rl_ptg_b200/csrc/ptg_capi.cu
echo variants/$name.so
