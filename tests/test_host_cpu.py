"""CPU tests (no GPU): the C-ABI library loads and exports what the header declares, the drop-in classes expose the
reference's surface, and the host-side helpers behave."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_in_header():
    from d3dp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "d3dp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(d3dp_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert sorted(_lib.SYMBOLS) == declared, "bindings out of sync with the header"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.d3dp_version()


def test_c_schedule_matches_reference_buffer():
    """The schedule d3dp_create computes for a C-ABI user who never calls d3dp_set_schedule (d3dp_schedule_host, plain
    C doubles with libm's cos) against the buffer the reference registers (common/diffusionpose.py:42-52,92-95; the
    oracle's schedule_buffers is bit-equal to it, oracle/validate_against_reference.py).  torch.cos is a vectorised
    SLEEF kernel, libm's is scalar: both are <= 1 ulp routines, and the cumulative product of 1000 factors may drift
    by a few ulp — the bound asserted here is 1e-13 relative (float32 coefficients need 6e-8); the Python surface
    always uploads torch's own buffers (D3DP._engine -> d3dp_set_schedule), which are the reference's bit for bit."""
    import ctypes as C

    import numpy as np

    from d3dp_b200 import _lib
    from oracle import d3dp_oracle as orc
    lib = _lib.load()
    out = np.empty(1000, dtype=np.float64)
    assert lib.d3dp_schedule_host(1000, out.ctypes.data_as(C.c_void_p)) == 0
    ref = orc.schedule_buffers(1000)["alphas_cumprod"].numpy()
    rel = np.abs(out - ref) / ref
    print(f"\n[schedule] C vs reference alphas_cumprod: bit-equal {int((out == ref).sum())}/1000, "
          f"max rel diff {rel.max():.2e}")
    assert rel.max() < 1e-13
    assert lib.d3dp_schedule_host(0, out.ctypes.data_as(C.c_void_p)) != 0


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from d3dp_b200._lib import D3dpError
    from d3dp_b200.engine import Engine
    with pytest.raises(D3dpError):
        Engine(frames=27)


def test_d3dp_module_surface_and_state_dict():
    from d3dp_b200 import D3DP
    from tests.util import JL, JR, make_args
    m = D3DP(make_args(27), JL, JR, is_train=False, num_proposals=3, sampling_timesteps=5)
    sd = m.state_dict()
    assert len(sd) == 220  # 12 float64 schedule buffers + 208 pose_estimator tensors
    bufs = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
            "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
            "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
    assert list(sd.keys())[:12] == bufs
    assert all(sd[b].dtype == torch.float64 and sd[b].shape == (1000,) for b in bufs)
    for k in ["pose_estimator.Spatial_pos_embed", "pose_estimator.Temporal_pos_embed",
              "pose_estimator.Spatial_patch_to_embedding.weight", "pose_estimator.time_mlp.1.weight",
              "pose_estimator.time_mlp.3.bias", "pose_estimator.STEblocks.7.attn.qkv.weight",
              "pose_estimator.TTEblocks.0.mlp.fc2.bias", "pose_estimator.Spatial_norm.weight",
              "pose_estimator.Temporal_norm.bias", "pose_estimator.head.0.weight", "pose_estimator.head.1.bias"]:
        assert k in sd, k
    assert sd["pose_estimator.Temporal_pos_embed"].shape == (1, 27, 512)
    assert sum(v.numel() for k, v in sd.items() if k.startswith("pose_estimator.")) == 34724867
    assert m._time_list() == [999, 799, 599, 399, 199, -1]
    for attr in ("forward", "ddim_sample", "ddim_sample_flip", "q_sample", "predict_noise_from_start",
                 "prepare_targets", "pose_estimator"):
        assert hasattr(m, attr)
    # strict load of a DataParallel-style checkpoint (keys prefixed `module.`, main.py:630) works like the reference
    dp_state = {"module." + k: v for k, v in sd.items()}
    torch.nn.DataParallel(m).load_state_dict(dp_state, strict=True) if torch.cuda.is_available() else \
        m.load_state_dict({k[len("module."):]: v for k, v in dp_state.items()}, strict=True)
    with pytest.raises(RuntimeError):  # parameters on the CPU: there is no CPU path
        m.pose_estimator.engine()


def test_default_init_matches_reference_when_available():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present on this machine")
    from d3dp_b200 import D3DP
    from tests.util import JL, JR, make_args
    torch.manual_seed(0)
    mine = D3DP(make_args(9), JL, JR, is_train=False, num_proposals=1, sampling_timesteps=1).state_dict()
    dp = rh.import_reference()
    torch.manual_seed(0)
    ref = dp.D3DP(rh.make_args(9), JL, JR, is_train=False, num_proposals=1, sampling_timesteps=1).state_dict()
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert mine[k].dtype == ref[k].dtype and torch.equal(mine[k], ref[k]), k


def test_unsupported_configurations_raise():
    from d3dp_b200 import MixSTE2
    with pytest.raises(ValueError):
        MixSTE2(num_frame=27, embed_dim_ratio=256, depth=8)
    with pytest.raises(ValueError):
        MixSTE2(num_frame=385, embed_dim_ratio=512, depth=8)


def test_flip_permutation_and_synthetic_determinism():
    from d3dp_b200.engine import flip_permutation
    from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, flip_2d,
                                     synthetic_pose_estimator_state)
    perm = flip_permutation(JL, JR)
    assert sorted(perm) == list(range(17)) and [perm[perm[j]] for j in range(17)] == list(range(17))
    x = torch.randn(2, 5, 17, 2)
    ref = x.clone()
    ref[..., 0] *= -1
    assert torch.equal(flip_2d(x), ref[:, :, perm])
    a, b = synthetic_pose_estimator_state(9, depth=1, seed=3), synthetic_pose_estimator_state(9, depth=1, seed=3)
    assert all(torch.equal(a[k], b[k]) for k in a) and len(a) == 8 + 24 + 8
    assert abs(a["STEblocks.0.attn.qkv.weight"].double().sum().item() - (-1.1920928955078125e-07)) < 50  # stable draw
    c = synthetic_pose_estimator_state(9, depth=1, seed=4)
    assert not torch.equal(a["head.1.weight"], c["head.1.weight"])


def test_shard_range_partitions_hypotheses():
    from d3dp_b200.distributed import shard_range
    for H in (1, 5, 20, 21, 160):
        for W in (1, 2, 3, 4, 8):
            parts = [shard_range(H, W, r) for r in range(W)]
            assert parts[0][0] == 0 and sum(n for _, n in parts) == H
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(W - 1))
            assert max(n for _, n in parts) - min(n for _, n in parts) <= 1


def test_clip_pipeline_matches_reference_semantics():
    """d3dp_b200.clips.eval_data_prepare == the oracle restatement of main.py:267-299 (and the reference itself where
    its tree is present), for sequence lengths below, equal to, and above multiples of F."""
    from d3dp_b200.clips import batches, eval_data_prepare, flip_inputs, stitch_clips
    from d3dp_b200.synthetic import H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, flip_2d
    from oracle import d3dp_oracle as orc
    from oracle import ref_harness as rh
    F = 27
    ref_fn = None
    if rh.available():
        import importlib.util
        import sys as _sys
        rh.import_reference()
        # main.py is a script; load only its eval_data_prepare by exec'ing the function source
        src = open(os.path.join(rh.REF, "main.py")).read()
        start = src.index("def eval_data_prepare(")
        end = src.index("\n\n\n", start)
        ns = {"torch": torch}
        from einops import rearrange
        ns["rearrange"] = rearrange
        exec(src[start:end], ns)
        ref_fn = ns["eval_data_prepare"]
    for n in (5, 27, 28, 54, 100):
        g = torch.Generator().manual_seed(n)
        seq2d = torch.randn(1, n, 17, 2, generator=g)
        seq3d = torch.randn(1, n, 17, 3, generator=g)
        c2, c3 = eval_data_prepare(F, seq2d, seq3d)
        assert torch.equal(c2, orc.eval_data_prepare(F, seq2d[0]))
        assert torch.equal(c3, orc.eval_data_prepare(F, seq3d[0]))
        if ref_fn is not None and n >= 2:
            r2, r3 = ref_fn(F, seq2d, seq3d)
            assert torch.equal(c2, r2) and torch.equal(c3, r3), n
        assert torch.equal(stitch_clips(c2, n), seq2d[0])  # clips cover the sequence exactly once
    assert torch.equal(flip_inputs(seq2d, JL, JR), flip_2d(seq2d))
    assert [(s.start, s.stop) for s in batches(10, 4)] == [(0, 4), (4, 8), (8, 10)]


def test_3dhp_variant_surface():
    from d3dp_b200.diffusionpose_3dhp import D3DP as D3DP3
    from tests.util import JL, JR, make_args
    m = D3DP3(make_args(9, depth=1), JL, JR, is_train=False, num_proposals=1, sampling_timesteps=1)
    assert m.OUTPUT_SCALE == 1000.0 and m.pose_estimator._output_scale == 1000.0 and len(m.state_dict()) == 12 + 40


class _OracleEngine:
    """Stand-in for Engine in host-logic tests: same jpma_gt / pmpjpe surface, computed by the oracle on CPU."""
    device = torch.device("cpu")

    def __init__(self, frames):
        self.frames = frames

    def jpma_gt(self, preds, traj, cam, x2d, gt, root_joint=0, linear=False):
        from oracle import d3dp_oracle as orc
        jagg, idx, pagg, e2d = orc.jpma(preds, traj, cam, x2d, root_joint, linear)
        e = orc.jpma_errors(preds, gt, traj, cam, x2d, root_joint, linear)
        return {"jagg_pose": jagg, "jagg_idx": idx.int(), "pagg_pose": pagg, "e2d_min": e2d, "e3d": e["e3d"],
                "jbest_pose": e["jbest_pose"]}

    def pmpjpe(self, preds, gt, root_joint=0):
        from oracle import d3dp_oracle as orc
        P = preds.clone()
        if root_joint >= 0:
            P[..., root_joint, :] = 0
        g = gt[:, None, None] if preds.dim() == 6 else gt[:, None]
        return orc.procrustes_errors(P, g)


def test_sequence_evaluator_packing_keeps_reference_numbers():
    """evaluate_sequences: packed cross-sequence batches give the numbers of the reference's per-sequence loop
    (main.py:685-724: per batch errors weighted by B*F), including the batch-dependent P-Best."""
    import types
    from d3dp_b200.clips import eval_data_prepare, flip_inputs
    from d3dp_b200.evaluate import evaluate_sequences
    from d3dp_b200.synthetic import H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR
    from oracle import d3dp_oracle as orc
    F, K, H, bs = 9, 2, 4, 3
    eng = _OracleEngine(F)
    model = types.SimpleNamespace(frames=F, pose_estimator=types.SimpleNamespace(engine=lambda: eng))
    g = torch.Generator().manual_seed(0)
    seqs = []
    for n in (5, 31, 40, 9):
        gt = 0.3 * torch.randn(n, 17, 3, generator=g) + torch.tensor([0.0, 0.0, 4.0])
        x2d = 0.3 * torch.randn(n, 17, 2, generator=g)
        seqs.append({"x2d": x2d, "gt": gt, "cam": torch.tensor([1.1, 1.1, 0.01, -0.02, -0.2, 0.1, 0.0, 0.001, -0.001])})

    def fake_preds(xb, fb):  # deterministic per clip, hypothesis- and step-dependent
        B = xb.shape[0]
        base = torch.cat([xb, xb[..., :1] * 0.5], dim=-1)[:, None, None]                     # [B,1,1,F,17,3]
        hk = torch.arange(K).reshape(1, K, 1, 1, 1, 1) * 0.01 + torch.arange(H).reshape(1, 1, H, 1, 1, 1) * 0.02
        return base + hk * (1 + fb[..., :1][:, None, None])

    kw = dict(kps_left=JL, kps_right=JR, batch_size=bs, protocol2=True, sampler=lambda xb, fb, bi: fake_preds(xb, fb))
    packed = evaluate_sequences(model, seqs, packed=True, return_poses=True, **kw)
    plain = evaluate_sequences(model, seqs, packed=False, **kw)
    assert packed["n_clips"] == 1 + 4 + 5 + 1 and packed["n_batches"] == 4 and plain["n_batches"] == 1 + 2 + 2 + 1
    # the reference loop, restated: per sequence, per batch of <= bs clips, errors weighted by B*F
    tot = {k: torch.zeros(K, dtype=torch.float64) for k in ("J-Best", "P-Best", "P-Agg", "J-Agg")}
    tot2 = {k: torch.zeros(K, dtype=torch.float64) for k in tot}
    N = 0
    for s in seqs:
        a, gt_c = eval_data_prepare(F, s["x2d"][None], s["gt"][None])
        b, _ = eval_data_prepare(F, flip_inputs(s["x2d"][None], JL, JR))
        traj = gt_c[:, :, :1].clone()
        gt_c = gt_c.clone()
        gt_c[:, :, 0] = 0
        for i in range(0, a.shape[0], bs):
            xb, fb, gb, tb = a[i:i + bs], b[i:i + bs], gt_c[i:i + bs], traj[i:i + bs]
            preds = fake_preds(xb, fb)
            cam = s["cam"][None].expand(xb.shape[0], 9)
            e = orc.jpma_errors(preds, gb, tb, cam, xb)
            _, idx, _, _ = orc.jpma(preds, tb, cam, xb)
            e2 = orc.p_jpma_errors(preds, gb, idx)
            w = xb.shape[0] * F
            for k in tot:
                tot[k] += w * e[k].double()
                tot2[k] += w * e2[k].double()
            N += w
    for k in tot:
        for res in (packed, plain):
            assert torch.allclose(res[k].double(), tot[k] / N, atol=1e-6), k
            assert torch.allclose(res["P2-" + k].double(), tot2[k] / N, atol=1e-6), k
    assert [p.shape[1] for p in packed["jagg_pose"]] == [5, 31, 40, 9] and packed["pagg_pose"][1].shape == (K, 31, 17, 3)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the reference's CPU path, oracle port) runs without a GPU and stdout carries exactly
    one JSON line with the contract's keys; everything else (library banners, warnings) is kept off stdout."""
    import json
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "poses/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gelu_polynomial_in_kernel_source_meets_its_accuracy_claim():
    """The fc1 epilogue's erf-GELU (gelu_erf_x2 in gemm_tcgen05.cuh: max(v,0) + w 2^q(max(w,-5.5)), w = -|v|) is
    re-evaluated here in float32 from the coefficients parsed out of the kernel source and compared with the float64
    erf form (reference: nn.GELU(), common/mixste.py:24,39): |error| <= 1e-6 everywhere and at most 0.02 ulp of the
    fp16 value the kernel stores."""
    import numpy as np
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "d3dp_b200", "csrc",
                            "gemm_tcgen05.cuh")).read()
    body = src[src.index("void gelu_erf_x2"):]
    body = body[:body.index("unpack_f32x2(g, a, b)")]
    coef = [np.float32(c) for c in re.findall(r"dup_f32x2\((-?[0-9.]+e[-+][0-9]+)f\)", body)]
    assert len(coef) == 7 and "fmaxf(wa, -5.5f)" in body
    f32 = np.float32
    v = np.concatenate([np.linspace(-9, 9, 400001), np.random.default_rng(0).normal(size=400000)]).astype(f32)
    w = np.minimum(v, -v)
    u = np.maximum(w, f32(-5.5))
    q = (u * coef[0] + coef[1]).astype(f32)                 # Horner in the kernel's order
    for c in coef[2:]:
        q = (q * u + c).astype(f32)
    e = np.exp2(q.astype(np.float64)).astype(f32)
    got = (w * e + np.maximum(v, f32(0))).astype(f32).astype(np.float64)
    vd = v.astype(np.float64)
    ref = 0.5 * vd * (1.0 + torch.erf(torch.from_numpy(vd) / np.sqrt(2.0)).numpy())
    err = np.abs(got - ref)
    ulp16 = np.maximum(2.0 ** -24, 2.0 ** (np.floor(np.log2(np.maximum(np.abs(ref), 1e-30))) - 10))
    assert err.max() <= 1e-6
    inside = np.abs(vd) <= 5.5                               # the fitted range; beyond it the result is -|v| 1.9e-8
    assert (err / ulp16)[inside].max() <= 0.02, (err / ulp16)[inside].max()
    assert err[~inside].max() <= 2e-7


def _ref_module(name):
    """Import a module of the reference tree where it exists (build container); None elsewhere (GPU box)."""
    from oracle import ref_harness as rh
    if not rh.available():
        return None
    import importlib
    import sys
    import types
    rh.import_reference()
    if "matplotlib" not in sys.modules:  # not installed; common/loss.py:1 imports one unused symbol from it
        sys.modules["matplotlib"] = types.ModuleType("matplotlib")
        mp = types.ModuleType("matplotlib.pyplot")
        mp.bone = None
        sys.modules["matplotlib.pyplot"] = mp
    return importlib.import_module(name)


@pytest.mark.parametrize("n_frames", [5, 27, 40, 54, 100])
def test_3dhp_export_layout_and_stitching(n_frames):
    """clips.export_layout_3dhp / stitch_clips_last_wins == the oracle restatement of main_3dhp.py:327-332,866-871
    (last clip overwrites the last F frames; MATLAB layout [3,17,N,K]), for sequences shorter than, equal to and not
    a multiple of F; the oracle restatement is itself compared with an exec of the reference's function."""
    import numpy as np

    from d3dp_b200.clips import eval_data_prepare, export_layout_3dhp, stitch_clips_last_wins
    from oracle import d3dp_oracle as orc
    F, K = 27, 3
    g = torch.Generator().manual_seed(n_frames)
    n_clips = max((n_frames + F - 1) // F, 1)
    clip_poses = torch.randn(n_clips, K, F, 17, 3, generator=g)
    if n_frames < F:   # what the pipeline produces for a padded short sequence: the real frames come first
        seq = torch.randn(K, n_frames, 17, 3, generator=g)
        clip_poses = torch.cat([seq, seq[:, -1:].expand(K, F - n_frames, 17, 3)], dim=1)[None]
        want = seq.permute(3, 2, 1, 0).numpy()
    else:
        want = orc.pose_post_process(clip_poses.numpy(), n_frames, F)
        # main_3dhp.py runs argparse and builds datasets at import: exec its function's source where the tree exists
        src_path = "/root/reference/main_3dhp.py"
        if os.path.exists(src_path):
            src = open(src_path).read()
            fn_src = src[src.index("def pose_post_process"):src.index("def cam_mm_to_pix")]
            ns = {}
            exec(fn_src, ns)
            dl = ns["pose_post_process"](clip_poses.numpy(), {"k": np.zeros((K, n_frames, 17, 3))}, "k", F)
            assert np.array_equal(dl["k"], want)
    out = export_layout_3dhp(clip_poses, n_frames)
    assert out.shape == (3, 17, n_frames, K)
    assert np.allclose(out.numpy(), want)
    st = stitch_clips_last_wins(clip_poses, n_frames)
    assert st.shape == (K, n_frames, 17, 3) and np.allclose(st.permute(3, 2, 1, 0).numpy(), want)
    # round trip with the clip cutter: cut -> identity per clip -> stitch gives the sequence back
    seq = torch.randn(n_frames, 17, 3, generator=g)
    cl, _ = eval_data_prepare(F, seq)
    assert torch.equal(stitch_clips_last_wins(cl[:, None], n_frames)[0], seq)


def test_valid_frame_metrics_pbest_pose_and_image_coordinates():
    """metrics.valid_frame_metrics (common/loss.py:109-145), metrics.pbest_pose (main_3dhp.py:785-795) and
    clips.image_coordinates (common/camera.py:14-18) against the oracle restatements, and those against the
    reference's own functions where its tree is present."""
    from d3dp_b200.clips import image_coordinates
    from d3dp_b200.metrics import pbest_pose, valid_frame_metrics
    from oracle import d3dp_oracle as orc
    g = torch.Generator().manual_seed(2)
    B, K, H, F = 3, 2, 4, 9
    gt = torch.randn(B, F, 17, 3, generator=g)
    gt[:, :, 14] = 0
    preds = gt[:, None, None] + 0.1 * torch.randn(B, K, H, F, 17, 3, generator=g)
    preds[:, :, :, :, 14] = 0
    valid = torch.rand(B, F, 1, generator=g) > 0.3
    e3d = torch.norm(preds - gt[:, None, None], dim=-1)
    pb, pa = valid_frame_metrics(e3d, preds.mean(dim=2), gt, valid)
    want_pb, want_pa = orc.mpjpe_3dhp_valid(preds, gt, valid), orc.mpjpe_3dhp_valid(preds, gt, valid, mean_pos=True)
    assert torch.allclose(pb, want_pb, atol=1e-6) and torch.allclose(pa, want_pa, atol=1e-6)
    pose, idx = orc.pbest_pose(preds, gt)
    mine = pbest_pose(preds, idx, root_joint=14)
    assert torch.equal(mine, pose)
    x = torch.randn(5, 17, 2, generator=g)
    assert torch.allclose(image_coordinates(x, 2048, 1536), orc.image_coordinates(x, 2048, 1536))
    loss = _ref_module("common.loss")
    if loss is not None:
        assert torch.allclose(loss.mpjpe_diffusion_3dhp(preds, gt, valid), want_pb, atol=1e-7)
        assert torch.allclose(loss.mpjpe_diffusion_3dhp(preds, gt, valid, mean_pos=True), want_pa, atol=1e-7)
        cam = _ref_module("common.camera")
        assert torch.allclose(torch.as_tensor(cam.image_coordinates(x.numpy(), 2048, 1536)).float(),
                              orc.image_coordinates(x, 2048, 1536), atol=1e-4)
    # no valid frame at all: NaN, like the reference's mean over an empty selection
    nb, na = valid_frame_metrics(e3d, preds.mean(dim=2), gt, torch.zeros(B, F, 1, dtype=torch.bool))
    assert torch.isnan(nb).all() and torch.isnan(na).all()
