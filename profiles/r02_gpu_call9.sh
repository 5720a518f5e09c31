#!/bin/bash
# Round 2, GPU call 9: final single-GPU validation of the tree (suite, smoke, bench line, launch list).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-250 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
D3DP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python profiles/launch_summary.py gpurun_out/launches.csv > gpurun_out/r02_launch_list_summary.txt 2>&1; cat gpurun_out/r02_launch_list_summary.txt
