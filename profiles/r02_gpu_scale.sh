#!/bin/bash
# Round 2: bench.py at N = 1, 2, 4, 8 back to back on ONE 8-GPU box (the driver's SCALE procedure), for the scaling table.
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
    print("N=$n value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "sampler min/max",
          round(d["phase_ms"]["sampler"]["min"], 1), round(d["phase_ms"]["sampler"]["max"], 1), "shard_check", d.get("shard_check"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("N=$n ERR", e)
PY
done | tee gpurun_out/r02_scale_samebox.txt
