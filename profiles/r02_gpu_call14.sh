#!/bin/bash
# Round 2, GPU call 14: embed kernel with fully coalesced 4-channel-per-quarter mapping: ncu time of the kernel in both
# builds, parity suite on the variant.
mkdir -p gpurun_out
L=$PWD/d3dp_b200/csrc
for lib in libd3dp_b200.so ab_embed.so; do
  D3DP_GRAPH=0 D3DP_LIB=$L/$lib timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none \
    -k regex:'embed_kernel|head_kernel|time_mlp' --csv --log-file gpurun_out/embed_$lib.csv python profiles/run_sampler.py 2 > /dev/null 2>&1
  echo $lib; grep -E "embed_kernel|head_kernel|time_mlp" gpurun_out/embed_$lib.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-200
done
D3DP_LIB=$L/ab_embed.so timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_embed.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_embed.log; tail -3 gpurun_out/pytest_embed.log
