#!/bin/bash
# Round 2, GPU call 6: LayerNorm-GEMM ring configurations at kernel level and inside the sampler (c3 width, K=1);
# initcheck location histogram.
mkdir -p gpurun_out
for v in a4b2r1 a3b2r2; do
  AB_ONLY=proj_res_ln,fc2_res_ln2,sampler AB_SAMPLER=4,20,1 AB_VISITS=2 timeout 400 python profiles/ab_lib.py ab_ln_a2b2r2.so ab_ln_$v.so > gpurun_out/ab_ln2_$v.log 2>&1; cat gpurun_out/ab_ln2_$v.log
done
export D3DP_GRAPH=0
SAN_FULL=0 timeout 500 compute-sanitizer --tool initcheck --show-backtrace no --print-limit 2000000 --kernel-name kns=d3dp \
  python profiles/sanitize_target.py 2>&1 | grep -E "^=========     at |ERROR SUMMARY|skipped" | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -30 > gpurun_out/sanitizer_initcheck_hist.txt
cat gpurun_out/sanitizer_initcheck_hist.txt
