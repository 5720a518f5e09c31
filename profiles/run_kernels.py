"""Launch each hot kernel a few times at the bench workload's shapes (c3 with flip: T = 2*4*20*17*243 rows) so that
ncu can capture them in isolation:  ncu --set full -k regex:<name> -c 1 python profiles/run_kernels.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_rooflines, measured_peaks  # noqa: E402
from d3dp_b200.engine import Engine  # noqa: E402

if __name__ == "__main__":
    n_streams = int(os.environ.get("D3DP_STREAMS", 2 * 4 * 20))
    eng = Engine(frames=243)
    T = n_streams * 17 * 243
    res = kernel_rooflines(eng, T, n_streams, measured_peaks())
    for k, v in res.items():
        print(k, {a: round(b, 4) for a, b in v.items()})
