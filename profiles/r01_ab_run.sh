#!/bin/bash
# gpurun --timeout 300 -- 'bash profiles/r01_ab_run.sh'
mkdir -p gpurun_out
timeout 280 python profiles/ab_attn.py > gpurun_out/ab_attn.log 2>&1; echo "rc=$?" >> gpurun_out/ab_attn.log
cat gpurun_out/ab_attn.log
