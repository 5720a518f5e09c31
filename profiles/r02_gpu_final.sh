#!/bin/bash
# Round 2, final GPU call: suite, smoke, ncu --set full of the six hot kernels as shipped (the summary bench.py
# parses), bench lines (c3 with CPU baseline, c2, reference arm), launch list of one bench step (graph replay).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
D3DP_PROFILE_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'gemm_2sm_kernel|attn_temporal_kernel|attn_spatial_kernel|gemm_ln_pair_kernel' -c 12 -f -o gpurun_out/r02_kernels_final \
  python profiles/run_kernels.py > gpurun_out/ncu_full.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_kernels_final.ncu-rep > gpurun_out/r02_ncu_kernels_summary.txt 2>&1; cut -c1-300 gpurun_out/r02_ncu_kernels_summary.txt
cp gpurun_out/r02_ncu_kernels_summary.txt profiles/r02_ncu_kernels_summary.txt
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-250 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --steps 10 --config c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 340 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches.csv > gpurun_out/r02_launch_list_summary.txt 2>&1; cat gpurun_out/r02_launch_list_summary.txt
