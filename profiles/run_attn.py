"""Launch the temporal-attention kernel at the bench shape a few times (for ncu): AB_LIB selects the library build."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from d3dp_b200 import _lib  # noqa: E402

if os.environ.get("AB_LIB"):
    _lib.LIB_PATH = os.environ["AB_LIB"]
from d3dp_b200.engine import Engine  # noqa: E402

eng = Engine(frames=243)
n_streams = 160
T = n_streams * 17 * 243
qkv = torch.randn(1024, 1536, generator=torch.Generator().manual_seed(0)).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    eng.test_attn(True, qkv, n_streams)
torch.cuda.synchronize()
