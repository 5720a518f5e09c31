"""Does it pay to keep an inter-kernel activation in L2?  Time per token (and board power) of the two attention kernels
when their 3 KB/token input is re-read from the 126 MB L2 (8 streams: 33 048 rows, 101 MB of qkv16, same buffer every
launch) against the bench shape (160 streams: 2 GB, always from HBM).  Sustained 1.5 s loops, nvidia-smi at 50 ms."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3dp_b200.engine import Engine  # noqa: E402
from profiles.power_profile import Smi  # noqa: E402

smi = Smi()
eng = Engine(frames=243)
g = torch.Generator().manual_seed(0)
print(f"{'kernel':16s} {'streams':>7s} {'rows':>8s} {'ms':>8s} {'ns/row':>7s} {'W':>6s} {'MHz':>6s} {'uJ/row':>7s}")
for n_streams in (8, 160):
    T = n_streams * 17 * 243
    base = torch.randn(1024, 1536, generator=g).half()
    qkv = base.repeat((T + 1023) // 1024, 1)[:T].contiguous().cuda()
    for name, temporal in (("attn_temporal", True), ("attn_spatial", False)):
        fn = lambda: eng.test_attn(temporal, qkv, n_streams)  # noqa: E731
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        s.record()
        reps = 0
        while time.time() - t0 < 1.5:
            for _ in range(100):
                fn()
            reps += 100
            torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        pw, clk, _ = smi.window(t0, time.time())
        print(f"{name:16s} {n_streams:7d} {T:8d} {ms:8.4f} {ms * 1e6 / T:7.2f} {pw:6.0f} {clk:6.0f} {pw * ms * 1e3 / T:7.3f}")
smi.p.terminate()
