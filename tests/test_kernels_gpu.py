"""Kernel-level parity on the GPU: each sm_100a kernel against a plain PyTorch fp32 restatement of the same op
computed on the CPU from the same fp16-rounded operands (so the only differences are accumulation order and the
fp16 rounding of the outputs)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _engine(F):
    from d3dp_b200.engine import Engine
    return Engine(frames=F)


def _ln(v, g, b, eps):
    return torch.nn.functional.layer_norm(v, (v.shape[-1],), g, b, eps)


@pytest.mark.parametrize("M", [128, 300, 128 * 150 + 17])
@pytest.mark.parametrize("mode", [0, 1])
def test_gemm_bias_modes(M, mode):
    eng = _engine(27)
    g = torch.Generator().manual_seed(M + mode)
    K, N = 512, (1536 if mode == 0 else 1024)
    a = torch.randn(M, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    bias = torch.randn(N, generator=g)
    ref = a.float() @ w.float().t() + bias
    if mode == 1:
        ref = torch.nn.functional.gelu(ref)
    out = eng.test_gemm(mode, a.cuda(), w.cuda(), bias.cuda()).float().cpu()
    err = (out - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item() / 8), err
    assert (out - ref).abs().mean().item() < 2e-3


def test_gemm_gelu_epilogue_is_exact_erf_to_fp16_rounding():
    """fc1's fused GELU (gelu_erf_x2: max(v,0) - |v| 2^q(|v|), degree-6 q) against float64 erf-GELU of the float64
    pre-activation: the only error left is the fp16 rounding of the stored value (rel 2^-11) plus fp32 accumulation
    noise, over pre-activations spanning +-8 (beyond the +-5.5 clamp of the fit) — mixste.py:24,39."""
    eng = _engine(27)
    g = torch.Generator().manual_seed(11)
    M, K, N = 600, 512, 1024
    a = torch.randn(M, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    bias = 2.5 * torch.randn(N, generator=g)
    pre = a.double() @ w.double().t() + bias.double()
    assert pre.min() < -7 and pre.max() > 7
    ref = 0.5 * pre * (1 + torch.erf(pre / math.sqrt(2)))
    out = eng.test_gemm(1, a.cuda(), w.cuda(), bias.cuda()).double().cpu()
    tol = 6e-4 * ref.abs() + 5e-6
    assert ((out - ref).abs() <= tol).all(), ((out - ref).abs() - tol).max().item()
    assert (out[pre < -6.5].abs() < 1e-6).all() and torch.allclose(out[pre > 6.5], pre[pre > 6.5], rtol=6e-4)


@pytest.mark.parametrize("M", [128, 27 * 17 * 3, 128 * 149 + 5])
@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("K", [512, 1024])
def test_gemm_layernorm_modes(M, mode, K):
    eng = _engine(27)
    F = 27
    g = torch.Generator().manual_seed(M * 7 + mode + K)
    a = torch.randn(M, K, generator=g).half()
    w = (torch.randn(512, K, generator=g) * 0.04).half()
    bias = torch.randn(512, generator=g) * 0.1
    x = torch.randn(M, 512, generator=g)
    ga, ba = 1 + 0.1 * torch.randn(512, generator=g), 0.1 * torch.randn(512, generator=g)
    gb, bb = 1 + 0.1 * torch.randn(512, generator=g), 0.1 * torch.randn(512, generator=g)
    tpos = 0.02 * torch.randn(F, 512, generator=g)
    v = x + a.float() @ w.float().t() + bias
    xd = x.cuda()
    if mode == 2:
        ref_x, ref_a = v, _ln(v, ga, ba, 1e-6)
        out = eng.test_gemm(2, a.cuda(), w.cuda(), bias.cuda(), x=xd, ln_a=(ga.cuda(), ba.cuda(), 1e-6))
    else:
        y = _ln(v, ga, ba, 1e-6) + tpos[torch.arange(M) % F]
        ref_x, ref_a = y, _ln(y, gb, bb, 1e-6)
        out = eng.test_gemm(3, a.cuda(), w.cuda(), bias.cuda(), x=xd, ln_a=(ga.cuda(), ba.cuda(), 1e-6),
                            ln_b=(gb.cuda(), bb.cuda(), 1e-6), tpos=tpos.cuda(), F=F)
    torch.cuda.synchronize()
    assert (xd.cpu() - ref_x).abs().max().item() < 2e-3
    assert (out.float().cpu() - ref_a).abs().max().item() < 1e-2


def _attn_ref(qkv, groups):
    """qkv [T,1536] fp32; groups: LongTensor [n_seq, L] of row indices forming each sequence."""
    T = qkv.shape[0]
    out = torch.zeros(T, 512)
    q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
    for h in range(8):
        sl = slice(64 * h, 64 * h + 64)
        Q, Kk, V = q[groups][..., sl], k[groups][..., sl], v[groups][..., sl]  # [n_seq, L, 64]
        att = torch.softmax(Q @ Kk.transpose(-1, -2) * 0.125, dim=-1)
        out[groups.reshape(-1), sl] = (att @ V).reshape(-1, 64)
    return out


@pytest.mark.parametrize("F,S", [(27, 3), (243, 2), (81, 2), (16, 1), (200, 1), (351, 1), (257, 1), (384, 1)])
def test_attention_temporal(F, S):
    eng = _engine(F)
    g = torch.Generator().manual_seed(F)
    T = S * 17 * F
    qkv = torch.randn(T, 1536, generator=g).half()
    groups = torch.arange(T).reshape(S * 17, F)
    ref = _attn_ref(qkv.float(), groups)
    out = eng.test_attn(True, qkv.cuda(), S).float().cpu()
    assert (out - ref).abs().max().item() < 8e-3


@pytest.mark.parametrize("F,S", [(27, 3), (243, 2), (9, 1)])
def test_attention_spatial(F, S):
    eng = _engine(F)
    g = torch.Generator().manual_seed(F + 1)
    T = S * 17 * F
    qkv = torch.randn(T, 1536, generator=g).half()
    # token order [S, 17, F]: sequence (s, f) = rows (s*17 + j)*F + f
    groups = torch.arange(T).reshape(S, 17, F).permute(0, 2, 1).reshape(S * F, 17)
    ref = _attn_ref(qkv.float(), groups)
    out = eng.test_attn(False, qkv.cuda(), S).float().cpu()
    assert (out - ref).abs().max().item() < 8e-3


def test_fp16_activations_saturate_instead_of_overflowing():
    """activations are stored as IEEE fp16; a value beyond +-65504 must clamp (F2FP.SATFINITE), not turn
    into inf (and NaN one softmax / LayerNorm later).  Bias of +-1e5 drives the qkv-style epilogue out of range."""
    eng = _engine(27)
    g = torch.Generator().manual_seed(3)
    M, K, N = 256, 512, 1536
    a = torch.randn(M, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * 0.05).half()
    bias = torch.zeros(N)
    bias[::2], bias[1::2] = 1e5, -1e5
    out = eng.test_gemm(0, a.cuda(), w.cuda(), bias.cuda()).float().cpu()
    assert torch.isfinite(out).all()
    assert (out[:, ::2] == 65504).all() and (out[:, 1::2] == -65504).all()
    gel = eng.test_gemm(1, a.cuda(), w[:1024].cuda(), bias[:1024].cuda()).float().cpu()
    # gelu(-1e5) = -1e5 * Phi(-1e5): the kernel clamps the argument of Phi at 5.5 (Phi = 1.9e-8), i.e. -0.0019 here
    assert torch.isfinite(gel).all() and (gel[:, ::2] == 65504).all() and (gel[:, 1::2].abs() < 3e-3).all()
