#!/bin/bash
# Round 2, GPU call 1: parity suite on the new code, TMEM micro-benchmark, A/B of the prepared switches, LayerNorm-GEMM
# ablations, ncu captures (launch list of the bench, --set full of the six hot kernels with source counters).
#   gpurun --timeout 1500 -- 'bash profiles/r02_gpu_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/smi.txt
timeout 60 profiles/ubench/tmem_bw > gpurun_out/tmem_bw.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py --steps 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
D3DP_GRAPH=0 timeout 200 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_nograph.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --config c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2>> gpurun_out/bench.err
# A/B: default vs the three prepared switches together, then the LayerNorm-GEMM ablations (mainloop only / epilogue only)
AB_TMP=/tmp timeout 500 python profiles/ab_lib.py libd3dp_b200.so ab_all3.so > gpurun_out/ab_all3.log 2>&1
AB_ONLY=proj_res_ln,fc2_res_ln2 AB_VISITS=1 timeout 200 python profiles/ab_lib.py ab_lnx1.so ab_lnx2.so > gpurun_out/ab_lnx.log 2>&1
# the variant build through the parity suite (kernel + sampler tests)
D3DP_LIB=$PWD/d3dp_b200/csrc/ab_all3.so timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_all3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_all3.log
# ncu: full set + source counters of the six hot kernels alone, then the launch list of a short bench
D3DP_PROFILE_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'gemm_2sm_kernel|attn_temporal_kernel|attn_spatial_kernel|gemm_ln_pair_kernel' -c 12 -f -o gpurun_out/r02_kernels \
  python profiles/run_kernels.py > gpurun_out/ncu_full.log 2>&1
D3DP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
cat gpurun_out/tmem_bw.txt; tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
cat gpurun_out/ab_all3.log gpurun_out/ab_lnx.log; tail -3 gpurun_out/pytest_all3.log
