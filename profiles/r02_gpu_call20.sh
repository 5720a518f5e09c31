#!/bin/bash
mkdir -p gpurun_out
timeout 200 python profiles/l2_residency_probe.py > gpurun_out/l2_residency_probe.txt 2>&1; cat gpurun_out/l2_residency_probe.txt
