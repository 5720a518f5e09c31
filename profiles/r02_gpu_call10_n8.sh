#!/bin/bash
# Round 2, 8-GPU call: BASELINE config 4 (bench at N=8: H=160 sharded 20/GPU, shard check) and the config-5 grid
# (subset) with hypothesis sharding over 8 GPUs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_n8.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?" >> gpurun_out/bench_n8.err
cut -c1-260 gpurun_out/bench_n8.json; tail -2 gpurun_out/bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n8.json").read())
    print("N=8 value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
          "phase", {k: v for k, v in d["phase_ms"].items() if k != "note"}, "shard_check", d.get("shard_check"))
except Exception as e:
    print("ERR", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 \
  profiles/sweep.py --quick > gpurun_out/sweep_n8.jsonl 2> gpurun_out/sweep_n8.err; echo "rc=$?" >> gpurun_out/sweep_n8.err
grep -v "^\*\|NCCL" gpurun_out/sweep_n8.jsonl; tail -2 gpurun_out/sweep_n8.err
