#!/bin/bash
mkdir -p gpurun_out
L=$PWD/d3dp_b200/csrc
D3DP_LIB=$L/ab_sp2.so timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_sp2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sp2.log; tail -3 gpurun_out/pytest_sp2.log
AB_ONLY=attn_spatial,sampler AB_SAMPLER=4,20,1 AB_VISITS=3 timeout 500 python profiles/ab_lib.py libd3dp_b200.so ab_sp2.so > gpurun_out/ab_sp2.log 2>&1; cat gpurun_out/ab_sp2.log
