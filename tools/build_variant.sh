#!/bin/bash
# build a variant of libomega4_cuda.so for A/B runs on one box: tools/build_variant.sh NAME [-D...]
# -> build/variants/NAME.so (git-ignored, travels with gpurun); select it with OMEGA4_CUDA_LIB=build/variants/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
  -o build/variants/$name.so audio-analyzer-omega_b200/csrc/omega4_cuda.cu
echo build/variants/$name.so
