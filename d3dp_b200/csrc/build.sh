#!/bin/bash
# Build libd3dp_b200.so for sm_100a (cross-compiles without a GPU).
#   D3DP_NVCC_EXTRA="-DD3DP_SMEM_PTRARITH=1" D3DP_OUT=ab_variant.so bash build.sh   # experiment build for profiles/ab_lib.py
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo \
  -Xcompiler -fPIC -shared ${D3DP_NVCC_EXTRA} \
  -o ${D3DP_OUT:-libd3dp_b200.so} d3dp_api.cu -lcudart_static -ldl -lrt -lpthread
