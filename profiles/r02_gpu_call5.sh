#!/bin/bash
# Round 2, GPU call 5: suite + smoke + bench lines on the new defaults, ncu --set full of the six hot kernels (the
# summary bench.py parses), launch list of one bench step, quick config-5 sweep.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "parity|passed|failed" gpurun_out/pytest_gpu.log | tail -24
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
D3DP_PROFILE_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:'gemm_2sm_kernel|attn_temporal_kernel|attn_spatial_kernel|gemm_ln_pair_kernel' -c 12 -f -o gpurun_out/r02_kernels \
  python profiles/run_kernels.py > gpurun_out/ncu_full.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_kernels.ncu-rep > gpurun_out/r02_ncu_kernels_summary.txt 2>&1; cut -c1-330 gpurun_out/r02_ncu_kernels_summary.txt
cp gpurun_out/r02_ncu_kernels_summary.txt profiles/r02_ncu_kernels_summary.txt
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-250 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --config c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2>> gpurun_out/bench.err; cut -c1-250 gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cut -c1-250 gpurun_out/bench_reference.json
D3DP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches.csv > gpurun_out/r02_launch_list_summary.txt 2>&1; cat gpurun_out/r02_launch_list_summary.txt
timeout 400 python profiles/sweep.py --quick > gpurun_out/sweep_quick.jsonl 2>&1; cat gpurun_out/sweep_quick.jsonl
