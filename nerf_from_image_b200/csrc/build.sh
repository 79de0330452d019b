#!/bin/bash
# Builds libnfi_render.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -O3 -std=c++17 --fmad=false -lineinfo \
  -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static \
  -I../../include ${NFI_PTXAS_V:+-Xptxas -v} \
  -o libnfi_render.so nfi_render.cu "$@"
