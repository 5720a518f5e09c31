import sys, torch
sys.path.insert(0, '.')
from d3dp_b200.engine import Engine
eng = Engine(frames=27)
torch.manual_seed(0)
for M in (128*3, 128*75, 128*149+5):
    K = 512
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g).half(); w = (torch.randn(512, K, generator=g)*0.04).half()
    bias = torch.randn(512, generator=g)*0.1; x = torch.randn(M, 512, generator=g)
    ga, ba = torch.ones(512), torch.zeros(512)
    v = x + a.float() @ w.float().t() + bias
    xd = x.cuda()
    out = eng.test_gemm(2, a.cuda(), w.cuda(), bias.cuda(), x=xd, ln_a=(ga.cuda(), ba.cuda(), 1e-6))
    torch.cuda.synchronize()
    err = (xd.cpu() - v).abs()
    T = (M + 127)//128
    pad = T*128 - M
    e = torch.cat([err, torch.zeros(pad, 512)]).reshape(T, 128, 4, 128).amax(dim=(1, 3))  # [tile, quarter]
    bad = (e > 2e-3).nonzero()
    print("M", M, "tiles", T, "max err", err.max().item(), "bad (tile,quarter) count", len(bad), bad[:20].tolist())
    ref_a = torch.nn.functional.layer_norm(v, (512,), ga, ba, 1e-6)
    ea = (out.float().cpu() - ref_a).abs()
    e2 = torch.cat([ea, torch.zeros(pad, 512)]).reshape(T, 128, 4, 128).amax(dim=(1, 3))
    bad2 = (e2 > 1e-2).nonzero()
    print("   a16 max err", ea.max().item(), "bad", len(bad2), bad2[:20].tolist())
