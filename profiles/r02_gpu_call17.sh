#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/power_profile.py > gpurun_out/power_profile.txt 2>&1; cat gpurun_out/power_profile.txt
