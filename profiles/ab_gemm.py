"""Same-box A/B of the cta_group::2 GEMM variants (D3DP_GEMM_DEEP=0/1) at the bench shape (T = 660 960 rows):
CUDA-event time of qkv (N=1536) and fc1+GELU (N=1024), and bit-equality of the outputs between the variants."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from d3dp_b200.engine import Engine  # noqa: E402

eng = Engine(frames=243)
T = 160 * 17 * 243
g = torch.Generator().manual_seed(0)
a = torch.randn(1024, 512, generator=g).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
outs = {}
for rnd in range(4):
    for deep in ("0", "2", "3"):
        os.environ["D3DP_GEMM_DEEP"] = deep
        for mode, N in ((0, 1536), (1, 1024)):
            gw = torch.Generator().manual_seed(N)
            w = (torch.randn(N, 512, generator=gw) * 0.03).half().cuda()
            bias = torch.randn(N, generator=gw).cuda()
            out = eng.test_gemm(mode, a, w, bias, F=243)
            torch.cuda.synchronize()
            key = (mode,)
            same = None
            if key in outs:
                same = torch.equal(out, outs[key])
            else:
                outs[key] = out.clone()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                eng.test_gemm(mode, a, w, bias, F=243)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 20
            print(f"deep={deep} mode={mode} N={N}: {ms:.4f} ms  {2.0 * T * N * 512 / ms / 1e9:.0f} TFLOP/s  equal-to-first: {same}",
                  flush=True)
