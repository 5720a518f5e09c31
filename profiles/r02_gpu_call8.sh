#!/bin/bash
# Round 2, GPU call 8: fine-grained epilogue of the cta_group::2 GEMM (32-column sub-slabs, 64-byte swizzle).
mkdir -p gpurun_out
L=$PWD/d3dp_b200/csrc
D3DP_LIB=$L/ab_fine3.so timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "bias_modes or gelu or saturate" > gpurun_out/pytest_fine.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fine.log; tail -3 gpurun_out/pytest_fine.log
D3DP_LIB=$L/ab_fine3.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_fine_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_fine_parity.log; tail -3 gpurun_out/pytest_fine_parity.log
AB_ONLY=qkv,fc1_gelu,sampler AB_SAMPLER=4,20,1 AB_VISITS=2 timeout 400 python profiles/ab_lib.py libd3dp_b200.so ab_fine3.so > gpurun_out/ab_fine3.log 2>&1; cat gpurun_out/ab_fine3.log
AB_ONLY=qkv,fc1_gelu AB_VISITS=1 timeout 300 python profiles/ab_lib.py ab_fine1.so ab_fine2.so > gpurun_out/ab_fine12.log 2>&1; cat gpurun_out/ab_fine12.log
