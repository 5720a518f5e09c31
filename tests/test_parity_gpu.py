"""End-to-end parity of the CUDA sampler (through the reference-shaped D3DP class -> C ABI) against
  (1) golden outputs produced by the unmodified reference (tests/golden/*.pt, made by tests/golden/make_golden.py),
  (2) the CPU oracle on fresh seeds,
with identical synthetic weights, 2-D keypoints and injected noise.  Tolerance (BASELINE.json north_star):
mean per-joint distance <= 1e-3 (fp32-class) — the fp16-operand / fp32-accumulate kernels are held to that tighter
budget, not the 5e-3 bf16 one.  Max per-joint distance is bounded at 1e-2 (clamp edges amplify single joints)."""
import pytest
import torch

from tests.util import GOLDEN_CASES, JL, JR, build_model, case_inputs, load_golden, mpjpe_distance

pytestmark = pytest.mark.gpu
MEAN_TOL, MAX_TOL = 1e-3, 1e-2


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sampler_matches_reference_golden(name):
    case = load_golden(name)
    sd, x2d, x2d_flip, n0, ns = case_inputs(case)
    model = build_model(case["F"], case["H"], case["K"], sd, case["scale"], case["depth"], case["flip"])
    if case["flip"]:
        out = model(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda()) if False else \
            model.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), noise_init=n0, noise_steps=ns)
    else:
        out = torch.stack(model.ddim_sample(x2d.cuda(), None, noise_init=n0, noise_steps=ns), dim=1)
    assert out.shape == case["preds"].shape and out.dtype == torch.float32
    mean, mx = mpjpe_distance(out, case["preds"])
    last_mean, _ = mpjpe_distance(out[:, -1], case["preds"][:, -1])
    print(f"\n[parity] {name}: mean {mean:.3e} max {mx:.3e} last-step mean {last_mean:.3e}")
    assert mean <= MEAN_TOL and mx <= MAX_TOL


def test_denoiser_forward_matches_reference_golden():
    case = load_golden("f27_flip")
    sd, x2d, _, n0, _ = case_inputs(case)
    model = build_model(case["F"], case["H"], case["K"], sd)
    out = model.pose_estimator(x2d.cuda(), n0.clamp(-1.1, 1.1).cuda(), case["denoise_t"].cuda())
    mean, mx = mpjpe_distance(out, case["denoise_out"])
    print(f"\n[parity] MixSTE2.forward: mean {mean:.3e} max {mx:.3e}")
    assert mean <= MEAN_TOL and mx <= MAX_TOL


@pytest.mark.parametrize("F,B,H,K,wseed", [(27, 1, 2, 2, 7), (81, 1, 1, 2, 3), (16, 2, 1, 3, 11), (351, 1, 1, 2, 5)])
def test_sampler_matches_oracle_fresh_seeds(F, B, H, K, wseed):
    from d3dp_b200.synthetic import synthetic_inputs, synthetic_pose_estimator_state
    from oracle import d3dp_oracle as orc
    sd = synthetic_pose_estimator_state(F, seed=wseed)
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H, K, F, seed=wseed + 100, noise_seed=wseed + 200)
    with torch.no_grad():
        ref = orc.ddim_sample(sd, x2d, x2d_flip, H, K, n0, ns, JL, JR)
    model = build_model(F, H, K, sd)
    out = model.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), noise_init=n0, noise_steps=ns)
    mean, mx = mpjpe_distance(out, ref)
    print(f"\n[parity] oracle F={F} B={B} H={H} K={K}: mean {mean:.3e} max {mx:.3e}")
    assert mean <= MEAN_TOL and mx <= MAX_TOL


def test_hypotheses_are_independent_bitwise():
    """Sharding property: running H=4 as two H=2 calls with the matching noise slices reproduces the
    single call bit for bit — the kernels are batch-invariant, so hypothesis sharding across GPUs is exact."""
    case = load_golden("f27_flip")
    sd, x2d, x2d_flip, _, _ = case_inputs(case)
    from d3dp_b200.synthetic import synthetic_inputs
    _, _, n0, ns = synthetic_inputs(2, 4, 3, 27)
    full = build_model(27, 4, 3, sd).ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(),
                                                      noise_init=n0, noise_steps=ns)
    half = build_model(27, 2, 3, sd)
    parts = [half.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), noise_init=n0[:, s],
                                   noise_steps=ns[:, :, s]) for s in (slice(0, 2), slice(2, 4))]
    assert torch.equal(full, torch.cat(parts, dim=2))


def test_sampler_stays_finite_and_close_with_large_weights():
    """parity was only shown at default-init weight scale.  Here the attention and MLP input projections of
    every block are scaled x4 (logits x16, hidden pre-activations x4 — the magnitude regime of a trained checkpoint
    with peaky attention heads): the fp16-operand pipeline must stay finite and inside the fp32-class budget."""
    from d3dp_b200.synthetic import synthetic_inputs, synthetic_pose_estimator_state
    from oracle import d3dp_oracle as orc
    F, B, H, K = 27, 2, 2, 3
    sd = synthetic_pose_estimator_state(F, seed=21)
    for k in sd:
        if k.endswith("attn.qkv.weight") or k.endswith("mlp.fc1.weight"):
            sd[k] = sd[k] * 4.0
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H, K, F, seed=31, noise_seed=41)
    with torch.no_grad():
        ref = orc.ddim_sample(sd, x2d, x2d_flip, H, K, n0, ns, JL, JR)
    out = build_model(F, H, K, sd).ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), noise_init=n0,
                                                   noise_steps=ns)
    assert torch.isfinite(out).all()
    mean, mx = mpjpe_distance(out, ref)
    print(f"\n[parity] x4 qkv/fc1 weights: mean {mean:.3e} max {mx:.3e}")
    assert mean <= MEAN_TOL and mx <= MAX_TOL


@pytest.mark.parametrize("F,B,H,K,flip", [(384, 1, 1, 1, True), (1, 3, 2, 2, True), (2, 1, 1, 2, False), (129, 1, 1, 1, True)])
def test_sampler_edge_sizes_match_oracle(F, B, H, K, flip):
    """Edge cases of the path: the largest supported clip (F = 384, three query tiles in the long temporal kernel), a
    single frame (temporal attention over one key), two frames without flip, and one frame past a 128-row query tile."""
    from d3dp_b200.synthetic import synthetic_inputs, synthetic_pose_estimator_state
    from oracle import d3dp_oracle as orc
    sd = synthetic_pose_estimator_state(F, seed=13)
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H, K, F, seed=F + 1, noise_seed=F + 2)
    with torch.no_grad():
        ref = orc.ddim_sample(sd, x2d, x2d_flip if flip else None, H, K, n0, ns, JL, JR)
    model = build_model(F, H, K, sd, flip=flip)
    if flip:
        out = model.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), noise_init=n0, noise_steps=ns)
    else:
        out = torch.stack(model.ddim_sample(x2d.cuda(), None, noise_init=n0, noise_steps=ns), dim=1)
    mean, mx = mpjpe_distance(out, ref)
    print(f"\n[parity] edge F={F} B={B} H={H} K={K} flip={flip}: mean {mean:.3e} max {mx:.3e}")
    assert out.shape == ref.shape and mean <= MEAN_TOL and mx <= MAX_TOL


def test_invalid_calls_fail_loudly():
    """Empty batches, a wrong frame count and a shard range outside H_total are refused with an error, not computed."""
    from d3dp_b200._lib import D3dpError
    case = load_golden("f27_flip")
    sd, x2d, x2d_flip, n0, ns = case_inputs(case)
    m = build_model(27, 2, 2, sd)
    with pytest.raises(D3dpError):
        m.ddim_sample_flip(x2d[:0].cuda(), None, input_2d_flip=x2d_flip[:0].cuda())
    with pytest.raises(D3dpError):
        m.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), seed=1, h_offset=-1, H_total=4)
    with pytest.raises(D3dpError):
        m.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), seed=1, h_offset=3, H_total=4)
    with pytest.raises((D3dpError, AssertionError, RuntimeError)):
        m.pose_estimator.engine().jpma(torch.zeros(2, 2, 2, 26, 17, 3), torch.zeros(2, 27, 1, 3), torch.zeros(2, 9),
                                       torch.zeros(2, 27, 17, 2))
