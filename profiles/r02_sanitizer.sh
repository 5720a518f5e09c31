#!/bin/bash
# compute-sanitizer over every d3dp kernel (SURVEY §5 / VERDICT r1 item 8): memcheck, racecheck, synccheck, initcheck.
# Kernel-by-kernel launches (D3DP_GRAPH=0) so that every report names the launch; depth-2 model (same kernels, 8x fewer
# launches).  Logs -> gpurun_out/sanitizer_<tool>.log (racecheck: full report summarised by racecheck_summary.py)
export D3DP_GRAPH=0
for tool in memcheck racecheck synccheck initcheck; do
  extra="--print-limit 30"
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all --print-limit 200000"
  timeout 900 compute-sanitizer --tool $tool $extra --kernel-name kns=d3dp \
    python profiles/sanitize_target.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_$tool.log
  if [ "$tool" = "racecheck" ]; then
    python profiles/racecheck_summary.py gpurun_out/sanitizer_racecheck.log > gpurun_out/sanitizer_racecheck_summary.txt 2>&1
    grep -E "^F=|SUMMARY|rc=" gpurun_out/sanitizer_racecheck.log > gpurun_out/sanitizer_racecheck_tail.txt
    rm -f gpurun_out/sanitizer_racecheck.log
    head -40 gpurun_out/sanitizer_racecheck_summary.txt
  else
    tail -4 gpurun_out/sanitizer_$tool.log
  fi
done
