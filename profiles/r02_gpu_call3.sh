#!/bin/bash
# Round 2, GPU call 3: row-LN kernel (gemm_ln_row.cuh) correctness + A/B, mainloop-only ablations, sanitizer re-run.
mkdir -p gpurun_out
L=$PWD/d3dp_b200/csrc
D3DP_LIB=$L/ab_lnrow.so timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "layernorm" > gpurun_out/pytest_lnrow.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lnrow.log; tail -3 gpurun_out/pytest_lnrow.log
D3DP_LIB=$L/ab_lnrow.so timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_aux_gpu.py -m gpu -x -q > gpurun_out/pytest_lnrow_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_lnrow_parity.log; tail -3 gpurun_out/pytest_lnrow_parity.log
AB_ONLY=proj_res_ln,fc2_res_ln2,fc2_tpos,sampler timeout 400 python profiles/ab_lib.py libd3dp_b200.so ab_lnrow.so > gpurun_out/ab_lnrow.log 2>&1; cat gpurun_out/ab_lnrow.log
AB_ONLY=proj_res_ln,fc2_res_ln2,fc2_tpos AB_VISITS=1 timeout 300 python profiles/ab_lib.py ab_lnrow.so ab_lnrow31.so > gpurun_out/ab_lnrow31.log 2>&1; cat gpurun_out/ab_lnrow31.log
AB_ONLY=proj_res_ln,fc2_res_ln2 AB_VISITS=1 timeout 200 python profiles/ab_lib.py ab_lnx1.so ab_lnx1_s3.so > gpurun_out/ab_lnx1.log 2>&1; cat gpurun_out/ab_lnx1.log
AB_ONLY=qkv,fc1_gelu AB_VISITS=1 timeout 200 python profiles/ab_lib.py libd3dp_b200.so ab_g2x1.so > gpurun_out/ab_g2x1.log 2>&1; cat gpurun_out/ab_g2x1.log
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
bash profiles/r02_sanitizer.sh
