#!/bin/bash
# Round 2, 2-GPU call: multi-GPU GPU tests (nn.DataParallel replicas), bench at N=2 under torchrun (shard check,
# per-phase times), N=1 on the same box for the efficiency.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_n2.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "dataparallel or invalidate" -s > gpurun_out/pytest_n2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_n2.log; tail -4 gpurun_out/pytest_n2.log
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_n1_samebox.json 2> gpurun_out/bench_n2.err; cut -c1-200 gpurun_out/bench_n1_samebox.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2>> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json | cut -c1-300; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1_samebox.json", "bench_n2.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read())
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "phase", {k: v for k, v in d["phase_ms"].items() if k != "note"}, "shard_check", d.get("shard_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
