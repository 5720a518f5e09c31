#!/bin/bash
# Round 2, GPU call 13: suite (new edge-size tests) + smoke + bench on the current tree.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "edge|passed|failed|Error" gpurun_out/pytest_gpu.log | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-220 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
