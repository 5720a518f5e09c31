#!/bin/bash
# Round 2, GPU call 16: same-box comparison of the round-1 tree (commit e749a89, rebuilt) and the current tree:
# bench.py --steps 3 --no-cpu-baseline, alternating, two visits each.
mkdir -p gpurun_out
R=$PWD
for v in 1 2; do
  (cd $R/profiles/_r1_tree && timeout 300 python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('round1 tree : value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'])")
  (cd $R && timeout 300 python bench.py --steps 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('current tree: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'])")
done | tee gpurun_out/r02_vs_r01_samebox.txt
