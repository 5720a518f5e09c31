"""Historical (round 1, session 2): produced r01b_ab_gemm.log (an earlier revision looped over the D3DP_GEMM_DEEP stage
variants inside one process).  Superseded by profiles/ab_lib.py.

Same-box A/B of two builds of the cta_group::2 GEMMs at the bench shape (T = 660 960 rows): CUDA-event time of qkv
(N=1536) and fc1+GELU (N=1024) and a checksum of the outputs (builds must agree bit for bit).
    python profiles/ab_gemm.py                  # driver: ab_prev.so, libd3dp_b200.so, alternating, own processes
    AB_LIB=path python profiles/ab_gemm.py one"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import torch
    from d3dp_b200 import _lib
    if os.environ.get("AB_LIB"):
        _lib.LIB_PATH = os.environ["AB_LIB"]
    from d3dp_b200.engine import Engine
    eng = Engine(frames=243)
    T = 160 * 17 * 243
    g = torch.Generator().manual_seed(0)
    a = torch.randn(1024, 512, generator=g).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
    res = []
    for mode, N in ((0, 1536), (1, 1024)):
        gw = torch.Generator().manual_seed(N)
        w = (torch.randn(N, 512, generator=gw) * 0.03).half().cuda()
        bias = torch.randn(N, generator=gw).cuda()
        out = eng.test_gemm(mode, a, w, bias, F=243)
        chk = out.view(torch.int16).to(torch.int64).sum().item()
        times = []
        for rnd in range(3):
            for _ in range(3):
                eng.test_gemm(mode, a, w, bias, F=243)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(20):
                eng.test_gemm(mode, a, w, bias, F=243)
            e.record()
            torch.cuda.synchronize()
            times.append(round(s.elapsed_time(e) / 20, 4))
        res.append(f"mode={mode} N={N}: ms {times} ({2.0 * T * N * 512 / min(times) / 1e9:.0f} TFLOP/s) checksum {chk}")
    print(f"lib={os.path.basename(os.environ.get('AB_LIB', 'default'))}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one()
    else:
        csrc = os.path.join(ROOT, "d3dp_b200", "csrc")
        for lib in ("ab_prev.so", "libd3dp_b200.so", "ab_prev.so", "libd3dp_b200.so"):
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"],
                           env=dict(os.environ, AB_LIB=os.path.join(csrc, lib)), timeout=120)
