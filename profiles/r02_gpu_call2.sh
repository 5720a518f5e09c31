#!/bin/bash
# Round 2, GPU call 2: bench lines (c3 graph / no graph, c2), A/B of kernel variants, compute-sanitizer.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
D3DP_GRAPH=0 timeout 200 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_nograph.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --config c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2>> gpurun_out/bench.err
D3DP_GRAPH=0 timeout 300 python bench.py --steps 5 --config c2 --no-cpu-baseline > gpurun_out/bench_c2_nograph.json 2>> gpurun_out/bench.err
AB_ONLY=proj_res_ln,fc2_res_ln2,fc2_tpos timeout 300 python profiles/ab_lib.py libd3dp_b200.so ab_ln31.so > gpurun_out/ab_ln31.log 2>&1
AB_ONLY=attn_temporal timeout 300 python profiles/ab_lib.py libd3dp_b200.so ab_attnsr.so > gpurun_out/ab_attnsr.log 2>&1
AB_ONLY=qkv,fc1_gelu,fc2_tpos,sampler timeout 300 python profiles/ab_lib.py libd3dp_b200.so ab_epipipe.so > gpurun_out/ab_epipipe.log 2>&1
D3DP_LIB=$PWD/d3dp_b200/csrc/ab_attnsr.so timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > gpurun_out/pytest_attnsr.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_attnsr.log
bash profiles/r02_sanitizer.sh
cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_nograph.json; cut -c1-300 gpurun_out/bench_c2.json; cut -c1-300 gpurun_out/bench_c2_nograph.json
cat gpurun_out/ab_ln31.log gpurun_out/ab_attnsr.log gpurun_out/ab_epipipe.log; tail -3 gpurun_out/pytest_attnsr.log
