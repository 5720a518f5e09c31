"""Historical (round 1, session 2): produced r01b_ab_attn.log by comparing the build saved as ab_prev.so with the
current one; the D3DP_ATTN_POLY knob it sets no longer exists.  Superseded by profiles/ab_lib.py.

Same-box A/B of temporal-attention builds / knobs: parity against the fp32 torch restatement and CUDA-event timing
at the bench shape (160 streams x 17 joints x 243).  Each configuration runs in its own process:
    python profiles/ab_attn.py                      # driver: loops over (library, D3DP_ATTN_POLY)
    AB_LIB=path python profiles/ab_attn.py one      # one configuration"""
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
    from tests.test_kernels_gpu import _attn_ref

    errs = []
    for F, S in ((243, 2), (200, 1), (129, 1), (128, 1), (27, 3), (16, 1)):
        eng = Engine(frames=F)
        g = torch.Generator().manual_seed(F)
        T = S * 17 * F
        qkv = torch.randn(T, 1536, generator=g).half()
        ref = _attn_ref(qkv.float(), torch.arange(T).reshape(S * 17, F))
        out = eng.test_attn(True, qkv.cuda(), S).float().cpu()
        errs.append(round((out - ref).abs().max().item(), 6))
    eng = Engine(frames=243)
    n_streams = 160
    T = n_streams * 17 * 243
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(1024, 1536, generator=g).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
    times = []
    for rnd in range(3):
        for _ in range(2):
            eng.test_attn(True, qkv, n_streams)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(10):
            eng.test_attn(True, qkv, n_streams)
        e.record()
        torch.cuda.synchronize()
        times.append(round(s.elapsed_time(e) / 10, 4))
    print(f"lib={os.path.basename(os.environ.get('AB_LIB', 'default'))} poly={os.environ.get('D3DP_ATTN_POLY', '0')}: "
          f"ms {times}  parity max-abs-err {errs}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one()
    else:
        csrc = os.path.join(ROOT, "d3dp_b200", "csrc")
        for lib, poly in (("ab_prev.so", "0"), ("libd3dp_b200.so", "0"), ("libd3dp_b200.so", "1"), ("ab_prev.so", "0")):
            env = dict(os.environ, AB_LIB=os.path.join(csrc, lib), D3DP_ATTN_POLY=poly)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, timeout=120)
