"""Power / clock / energy per launch of each hot kernel in a sustained loop (and of the whole sampler), from nvidia-smi
samples every 50 ms: under the 1 kW cap the step time follows the ENERGY of the kernels, not their stall-free time.
    python profiles/power_profile.py  > gpurun_out/power_profile.txt"""
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3dp_b200.engine import Engine  # noqa: E402


class Smi:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm,clocks.mem,temperature.gpu",
                                   "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.p.stdout:
            try:
                self.rows.append((time.time(),) + tuple(float(x) for x in line.split(",")))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 + 0.3 <= x[0] <= t1]
        n = max(len(r), 1)
        return sum(x[1] for x in r) / n, sum(x[2] for x in r) / n, len(r)


def main():
    smi = Smi()
    eng = Engine(frames=243)
    n_streams = 160
    T = n_streams * 17 * 243
    g = torch.Generator().manual_seed(0)
    a512 = torch.randn(1024, 512, generator=g).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
    a1024 = torch.cat([a512, a512], dim=1)
    qkv = torch.cat([a512, a512, a512], dim=1).contiguous()
    x = torch.zeros(T, 512, device="cuda")
    ones, zeros, bias = torch.ones(512, device="cuda"), torch.zeros(512, device="cuda"), torch.zeros(1536, device="cuda")
    w = lambda n, k: (torch.randn(n, k, generator=g) * 0.03).half().cuda()  # noqa: E731
    w_qkv, w_fc1, w_proj, w_fc2 = w(1536, 512), w(1024, 512), w(512, 512), w(512, 1024)
    jobs = [
        ("idle", None),
        ("gemm_qkv", lambda: eng.test_gemm(0, a512, w_qkv, bias, F=243)),
        ("gemm_fc1_gelu", lambda: eng.test_gemm(1, a512, w_fc1, bias, F=243)),
        ("gemm_proj_res_ln", lambda: eng.test_gemm(2, a512, w_proj, bias, x=x, ln_a=(ones, zeros, 1e-6), F=243)),
        ("gemm_fc2_res_ln2", lambda: eng.test_gemm(3, a1024, w_fc2, bias, x=x, ln_a=(ones, zeros, 1e-6),
                                                   ln_b=(ones, zeros, 1e-6), F=243)),
        ("attn_temporal", lambda: eng.test_attn(True, qkv, n_streams)),
        ("attn_spatial", lambda: eng.test_attn(False, qkv, n_streams)),
    ]
    print(f"{'kernel':18s} {'ms/launch':>9s} {'W':>7s} {'SM MHz':>7s} {'J/launch':>9s} {'launches/step':>13s} {'J/step':>8s}")
    per_step = {"gemm_qkv": 160, "gemm_fc1_gelu": 160, "gemm_proj_res_ln": 160, "gemm_fc2_res_ln2": 160,
                "attn_temporal": 80, "attn_spatial": 80}
    total = 0.0
    for name, fn in jobs:
        if fn is None:
            t0 = time.time()
            time.sleep(1.5)
            pw, clk, n = smi.window(t0, time.time())
            print(f"{name:18s} {'':>9s} {pw:7.0f} {clk:7.0f}")
            continue
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        s.record()
        reps = 0
        while time.time() - t0 < 2.0:
            for _ in range(50):
                fn()
            reps += 50
            torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize()
        t1 = time.time()
        ms = s.elapsed_time(e) / reps
        pw, clk, n = smi.window(t0, t1)
        j = pw * ms * 1e-3
        total += j * per_step[name]
        print(f"{name:18s} {ms:9.4f} {pw:7.0f} {clk:7.0f} {j:9.3f} {per_step[name]:13d} {j * per_step[name]:8.1f}")
    print(f"sum over the six kernels: {total:.0f} J per sampler call (c3)  -> {total / 1000.0 * 1e3:.0f} ms at a 1000 W cap")
    smi.p.terminate()


if __name__ == "__main__":
    main()
