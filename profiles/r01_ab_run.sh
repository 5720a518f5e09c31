#!/bin/bash
# gpurun --timeout 400 -- 'bash profiles/r01_ab_run.sh'
mkdir -p gpurun_out
timeout 300 python profiles/ab_gemm.py > gpurun_out/ab_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/ab_gemm.log
cat gpurun_out/ab_gemm.log
