#!/bin/bash
# Build libd3dp_b200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo \
  -Xcompiler -fPIC -shared ${D3DP_NVCC_EXTRA} \
  -o libd3dp_b200.so d3dp_api.cu -lcudart_static -ldl -lrt -lpthread
