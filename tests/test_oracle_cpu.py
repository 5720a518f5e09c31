"""CPU tests (no GPU): the oracle against the reference-generated golden vectors and known answers."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import d3dp_oracle as orc
from tests.util import GOLDEN_CASES, JL, JR, case_inputs, load_golden


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_reference_golden(name):
    """tests/golden/*.pt were produced by the unmodified reference (tests/golden/make_golden.py); the oracle is a
    restatement with the same ATen ops, so it must agree to float32 re-association noise (observed: bit-exact)."""
    case = load_golden(name)
    sd, x2d, x2d_flip, n0, ns = case_inputs(case)
    torch.set_num_threads(8)
    with torch.no_grad():
        out = orc.ddim_sample(sd, x2d, x2d_flip if case["flip"] else None, case["H"], case["K"], n0, ns, JL, JR,
                              scale=case["scale"], depth=case["depth"])
    mean, mx = orc.mpjpe_distance(out, case["preds"])
    assert out.shape == case["preds"].shape
    assert mx < 2e-5, (mean, mx)


def test_oracle_denoiser_golden():
    case = load_golden("f27_flip")
    sd, x2d, _, n0, _ = case_inputs(case)
    with torch.no_grad():
        out = orc.denoiser(sd, x2d, n0.clamp(-1.1, 1.1), case["denoise_t"])
    assert orc.mpjpe_distance(out, case["denoise_out"])[1] < 1e-5


def test_schedule_known_answers():
    """SURVEY Appendix A.1 values, probed from the reference's registered buffers."""
    b = orc.schedule_buffers(1000)
    ac = b["alphas_cumprod"]
    assert b["betas"].dtype == torch.float64
    np.testing.assert_allclose(b["betas"][0].item(), 4.128422482196914e-05, rtol=1e-12)
    assert b["betas"][999].item() == 0.999
    for t, v in [(999, 2.4287669070e-09), (899, 2.4091724140e-02), (799, 9.4045612677e-02), (599, 3.4080963976e-01),
                 (399, 6.4747821115e-01), (199, 8.9870592060e-01), (99, 9.7209273711e-01), (0, 9.9995871578e-01)]:
        np.testing.assert_allclose(ac[t].item(), v, rtol=2e-10)
    np.testing.assert_allclose(b["sqrt_recip_alphas_cumprod"][899].item(), 6.4426725552, rtol=1e-10)
    np.testing.assert_allclose(b["sqrt_recipm1_alphas_cumprod"][899].item(), 6.3645918685, rtol=1e-10)
    # DDIM step coefficients (eta = 1)
    for (t, tn), (sig, c, sq) in {(999, 899): (9.8788065061e-01, 3.0986175726e-04, 1.5521508992e-01),
                                   (199, 99): (1.4421876293e-01, 8.4310208785e-02, 9.8594763406e-01)}.items():
        a, an = ac[t], ac[tn]
        sigma = ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        cc = (1 - an - sigma ** 2).sqrt()
        np.testing.assert_allclose([sigma.item(), cc.item(), an.sqrt().item()], [sig, c, sq], rtol=2e-9)


def test_time_lists_known_answers():
    assert orc.time_list(1000, 1) == [999, -1]
    assert orc.time_list(1000, 2) == [999, 499, -1]
    assert orc.time_list(1000, 3) == [999, 665, 332, -1]
    assert orc.time_list(1000, 4) == [999, 749, 499, 249, -1]
    assert orc.time_list(1000, 5) == [999, 799, 599, 399, 199, -1]
    assert orc.time_list(1000, 10) == [999, 899, 799, 699, 599, 499, 399, 299, 199, 99, -1]


def test_c_time_list_matches_torch_linspace():
    """d3dp_time_list (host helper of the C ABI) restates torch.linspace(...).int() exactly for every K."""
    from d3dp_b200 import _lib
    lib = _lib.load()
    for K in list(range(1, 130)) + [200, 250, 333, 500, 999, 1000]:
        buf = (C.c_int32 * (K + 1))()
        assert lib.d3dp_time_list(1000, K, buf) == 0
        assert list(buf) == orc.time_list(1000, K), K


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = orc.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in out) == exp


def test_philox_normal_statistics():
    z = orc.philox_normal(seed=42, draw=3, elem=np.arange(200000, dtype=np.uint64))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert np.array_equal(z, orc.philox_normal(42, 3, np.arange(200000, dtype=np.uint64)))
    assert not np.allclose(z[:100], orc.philox_normal(43, 3, np.arange(100, dtype=np.uint64)))


def test_jpma_oracle_properties():
    from d3dp_b200.synthetic import synthetic_camera
    g = torch.Generator().manual_seed(0)
    B, K, H, F = 2, 3, 5, 7
    preds = 0.4 * torch.randn(B, K, H, F, 17, 3, generator=g)
    x2d = 0.3 * torch.randn(B, F, 17, 2, generator=g)
    traj, cam = synthetic_camera(B, F)
    jagg, idx, pagg, e2d = orc.jpma(preds, traj, cam, x2d)
    assert jagg.shape == (B, K, F, 17, 3) and idx.shape == (B, K, F, 17)
    assert torch.all(jagg[:, :, :, 0] == 0) and torch.all(pagg[:, :, :, 0] == 0)  # root joint zeroed
    # a hypothesis whose reprojection is exact must be selected
    P = preds.clone()
    P[:, :, :, :, 0] = 0
    X = P[:, :, 2] + traj.reshape(B, 1, F, 1, 3)
    uv = orc.project_to_2d(X.reshape(B, K * F, 17, 3), cam).reshape(B, K, F, 17, 2)
    _, idx2, _, e2 = orc.jpma(preds, traj, cam, uv[:, -1])  # 2-D input = reprojection of hypothesis 2 at the last step
    assert torch.all(idx2[:, -1, :, 1:] == 2) and e2[:, -1].max() < 1e-6
    # H = 1: J-Agg == P-Agg == the hypothesis
    j1, _, p1, _ = orc.jpma(preds[:, :, :1], traj, cam, x2d)
    assert torch.equal(j1, p1)


def test_torch_norm_association_is_the_fma_chain_the_jpma_kernel_uses():
    """The JPMA kernel takes its per-joint argmin over e2d = sqrt(fma(dv, dv, du*du)) and e3d = sqrt(fma(dz, dz,
    fma(dy, dy, dx*dx))) (d3dp_b200/csrc/elementwise.cuh).  That is exactly how ATen's CPU torch.norm(dim=-1)
    accumulates over a short last axis, which is what the reference's loss functions call (common/loss.py:54-76):
    checked here bit for bit with float64 emulation of the fused multiply-adds, so index-exact argmin parity on the
    GPU (tests/test_aux_gpu.py::test_jpma_matches_oracle) rests on a checked identity, not on luck."""
    g = torch.Generator().manual_seed(5)
    d = torch.randn(50000, 3, generator=g) * torch.logspace(-3, 1, 50000)[:, None]
    f32 = np.float32
    a = d.numpy().astype(np.float64)

    def rn(x):
        return x.astype(f32).astype(np.float64)
    e2 = np.sqrt(rn(a[:, 1] * a[:, 1] + rn(a[:, 0] * a[:, 0])).astype(f32))
    e3 = np.sqrt(rn(a[:, 2] * a[:, 2] + rn(a[:, 1] * a[:, 1] + rn(a[:, 0] * a[:, 0]))).astype(f32))
    assert np.array_equal(torch.norm(d[:, :2], dim=-1).numpy().view(np.uint32), e2.view(np.uint32))
    assert np.array_equal(torch.norm(d, dim=-1).numpy().view(np.uint32), e3.view(np.uint32))
