#!/bin/bash
# Round 2, GPU call 15: suite + bench c2 / c3 after the time-MLP batching and embed / head changes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --config c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_c2.json
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2>> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
