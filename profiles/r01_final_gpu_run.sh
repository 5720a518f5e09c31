#!/bin/bash
# One gpurun call: GPU test suite, smoke, bench line, ncu --set full of the kernels changed this session, short ncu
# launch list.   gpurun --timeout 900 -- 'bash profiles/r01_final_gpu_run.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 420 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
D3DP_PROFILE_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:'gemm_2sm_kernel|attn_temporal_kernel' -c 6 -f -o gpurun_out/r01b_kernels python profiles/run_kernels.py \
  > gpurun_out/ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
