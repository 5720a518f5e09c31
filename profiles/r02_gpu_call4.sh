#!/bin/bash
# Round 2, GPU call 4: split A/B operand rings in the LayerNorm GEMMs (variants), new default build through the suite,
# sanitizer racecheck/initcheck with aggregated reports.
mkdir -p gpurun_out
L=$PWD/d3dp_b200/csrc
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; grep -E "parity|passed|failed" gpurun_out/pytest_gpu.log | tail -22
for v in a4b2r1x0 a5b2r1x0 a4b2r1x1 a6b1r1x0; do
  AB_ONLY=proj_res_ln,fc2_res_ln2 AB_VISITS=1 timeout 200 python profiles/ab_lib.py ab_ln_a2b2r2x0.so ab_ln_$v.so > gpurun_out/ab_ln_$v.log 2>&1; cat gpurun_out/ab_ln_$v.log
done
timeout 400 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
export D3DP_GRAPH=0
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis --show-backtrace no --print-limit 400 --kernel-name kns=d3dp \
  python profiles/sanitize_target.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_racecheck.log
python profiles/racecheck_summary.py gpurun_out/sanitizer_racecheck.log > gpurun_out/sanitizer_racecheck_summary.txt 2>&1; head -50 gpurun_out/sanitizer_racecheck_summary.txt
SAN_FULL=0 timeout 400 compute-sanitizer --tool initcheck --show-backtrace no --print-limit 200 --kernel-name kns=d3dp \
  python profiles/sanitize_target.py > gpurun_out/sanitizer_initcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_initcheck.log
grep -E "Uninitialized|at .*cuh|at .*\.cu" gpurun_out/sanitizer_initcheck.log | sed -E 's/0x[0-9a-f]+//g; s/thread \([0-9,]+\)//; s/block \([0-9,]+\)//' | sort | uniq -c | sort -rn | head -20
tail -3 gpurun_out/sanitizer_initcheck.log
