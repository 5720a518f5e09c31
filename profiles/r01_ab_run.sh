#!/bin/bash
# Same-box A/B of a variant build against the default build (see profiles/README.md, "Experiment switches"):
#   D3DP_NVCC_EXTRA="-DD3DP_SMEM_PTRARITH=1" D3DP_OUT=ab_variant.so bash d3dp_b200/csrc/build.sh
#   gpurun --timeout 400 -- 'bash profiles/r01_ab_run.sh'
mkdir -p gpurun_out
timeout 300 python profiles/ab_lib.py > gpurun_out/ab_lib.log 2>&1; echo "rc=$?" >> gpurun_out/ab_lib.log
cat gpurun_out/ab_lib.log
