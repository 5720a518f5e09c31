"""Time the fc2 (EPI_RES_LN2) kernel with and without its second LayerNorm at the bench shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3dp_b200.engine import Engine

eng = Engine(frames=243)
T = 160 * 17 * 243
dev = eng.device
g = torch.Generator().manual_seed(0)
a = (torch.randn(1024, 1024, generator=g).half().repeat((T + 1023) // 1024, 1)[:T]).to(dev)
w = (torch.randn(512, 1024, generator=g) * 0.03).half().to(dev)
bias = torch.zeros(512, device=dev); ones = torch.ones(512, device=dev); zeros = torch.zeros(512, device=dev)
x = torch.zeros(T, 512, device=dev)


def timeit(fn, reps=8):
    for _ in range(2): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps

print("fc2 LN2 full      ", timeit(lambda: eng.test_gemm(3, a, w, bias, x=x, ln_a=(ones, zeros, 1e-6), ln_b=(ones, zeros, 1e-6), F=243)))
print("fc2 LN2 no ln_b   ", timeit(lambda: eng.test_gemm(3, a, w, bias, x=x, ln_a=(ones, zeros, 1e-6), F=243)))
a512 = a[:, :512].contiguous(); w512 = w[:, :512].contiguous()
print("K=512 LN2 full    ", timeit(lambda: eng.test_gemm(3, a512, w512, bias, x=x, ln_a=(ones, zeros, 1e-6), ln_b=(ones, zeros, 1e-6), F=243)))
print("K=512 LN (proj)   ", timeit(lambda: eng.test_gemm(2, a512, w512, bias, x=x, ln_a=(ones, zeros, 1e-6))))
print("K=1024 LN         ", timeit(lambda: eng.test_gemm(2, a, w, bias, x=x, ln_a=(ones, zeros, 1e-6))))
