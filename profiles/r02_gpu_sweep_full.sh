#!/bin/bash
# Round 2: the FULL BASELINE config-5 grid (36 cells) on one GPU.
mkdir -p gpurun_out
timeout 900 python profiles/sweep.py > gpurun_out/r02_sweep_full_1gpu.jsonl 2> gpurun_out/r02_sweep_full_1gpu.err
echo "rc=$?"; wc -l gpurun_out/r02_sweep_full_1gpu.jsonl; tail -3 gpurun_out/r02_sweep_full_1gpu.err
