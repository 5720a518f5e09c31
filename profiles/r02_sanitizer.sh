#!/bin/bash
# compute-sanitizer over every d3dp kernel (SURVEY §5 / VERDICT r1 item 8): memcheck, racecheck, synccheck, initcheck.
# Kernel-by-kernel launches (D3DP_GRAPH=0) so that every report names the launch; depth-2 model (same kernels, 8x fewer
# launches).  Logs -> gpurun_out/sanitizer_<tool>.log
export D3DP_GRAPH=0
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  [ "$tool" = "initcheck" ] && extra="--track-unused-memory no"
  timeout 600 compute-sanitizer --tool $tool $extra --kernel-regex kns=d3dp --print-limit 30 \
    python profiles/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
