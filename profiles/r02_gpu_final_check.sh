mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo bench rc=$?
python -c "
import json; d=json.loads([l for l in open('gpurun_out/final_bench.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['cpu_baseline'], d['gpu_launches'], d['clocks'])"
