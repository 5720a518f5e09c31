#!/bin/bash
# Round 2, GPU call 12: mbarrier try_wait with a suspend-time hint (fewer polling instructions under the power cap).
mkdir -p gpurun_out
for h in 1000 20000 1000000; do
  AB_ONLY=qkv,fc1_gelu,proj_res_ln,fc2_res_ln2,attn_temporal,sampler AB_SAMPLER=4,20,1 AB_VISITS=1 timeout 400 python profiles/ab_lib.py libd3dp_b200.so ab_hint$h.so > gpurun_out/ab_hint$h.log 2>&1; cat gpurun_out/ab_hint$h.log
done
