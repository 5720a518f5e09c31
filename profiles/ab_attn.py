"""Same-box A/B of the temporal-attention schedule knobs (D3DP_ATTN_LOCKSTEP, D3DP_ATTN_POLY): parity against the
fp32 torch restatement at F in {243, 200, 27} and CUDA-event timing at the bench shape (160 streams x 17 joints x 243)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3dp_b200.engine import Engine  # noqa: E402
from tests.test_kernels_gpu import _attn_ref  # noqa: E402


def parity(F, S):
    eng = Engine(frames=F)
    g = torch.Generator().manual_seed(F)
    T = S * 17 * F
    qkv = torch.randn(T, 1536, generator=g).half()
    ref = _attn_ref(qkv.float(), torch.arange(T).reshape(S * 17, F))
    out = eng.test_attn(True, qkv.cuda(), S).float().cpu()
    return (out - ref).abs().max().item()


def timing(reps=10):
    eng = Engine(frames=243)
    n_streams = 160
    T = n_streams * 17 * 243
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(1024, 1536, generator=g).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
    for _ in range(2):
        eng.test_attn(True, qkv, n_streams)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        eng.test_attn(True, qkv, n_streams)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


if __name__ == "__main__":
    for rnd in range(2):
        for lock, poly in ((1, 0), (0, 0), (0, 1), (1, 1)):
            os.environ["D3DP_ATTN_LOCKSTEP"], os.environ["D3DP_ATTN_POLY"] = str(lock), str(poly)
            errs = [parity(F, S) for F, S in ((243, 2), (200, 1), (27, 3))] if rnd == 0 else []
            print(f"lockstep={lock} poly={poly}: {timing():.4f} ms  parity max-abs-err {errs}", flush=True)
