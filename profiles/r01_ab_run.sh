#!/bin/bash
# gpurun --timeout 400 -- 'bash profiles/r01_ab_run.sh'
mkdir -p gpurun_out
timeout 300 python profiles/ab_slab.py > gpurun_out/ab_slab.log 2>&1; echo "rc=$?" >> gpurun_out/ab_slab.log
cat gpurun_out/ab_slab.log
