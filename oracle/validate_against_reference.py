"""Pin the oracle against the live reference (run in the build container, where /root/reference exists):
    python oracle/validate_against_reference.py
Checks the schedule buffers, the time lists, one denoiser forward, the full flip sampler and JPMA."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, synthetic_camera,  # noqa: E402
                                 synthetic_inputs, synthetic_pose_estimator_state)
from oracle import d3dp_oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402


def main():
    assert rh.available(), "reference tree not found"
    torch.set_num_threads(os.cpu_count())
    ok = True
    F, B, H, K = 27, 2, 3, 4
    sd = synthetic_pose_estimator_state(F, seed=0)
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H, K, F)
    model = rh.build_reference_model(F, H, K, sd, JL, JR)
    # schedule buffers: the oracle must reproduce the registered float64 buffers bit for bit
    bufs = orc.schedule_buffers(1000)
    for k, v in bufs.items():
        same = torch.equal(v, getattr(model, k))
        ok &= same
        print(f"buffer {k:34s} bit-equal: {same}")
    # one denoiser forward
    t = torch.full((B,), 499, dtype=torch.long)
    x_t = n0.clamp(-1.1, 1.1)
    with torch.no_grad():
        ref = model.pose_estimator(x2d, x_t, t)
        mine = orc.denoiser(sd, x2d, x_t, t)
    m, mx = orc.mpjpe_distance(mine, ref)
    print(f"denoiser forward   mean {m:.3e} max {mx:.3e}")
    ok &= mx < 1e-5
    # full sampler
    ref = rh.run_reference_sampler(model, x2d, x2d_flip, n0, ns)
    with torch.no_grad():
        mine = orc.ddim_sample(sd, x2d, x2d_flip, H, K, n0, ns, JL, JR)
    m, mx = orc.mpjpe_distance(mine, ref)
    print(f"ddim_sample_flip   mean {m:.3e} max {mx:.3e}  shape {tuple(ref.shape)}")
    ok &= mx < 2e-5
    # JPMA vs the reference's camera / loss functions
    sys.modules.setdefault("matplotlib", __import__("types").ModuleType("matplotlib"))
    import types
    mp = types.ModuleType("matplotlib.pyplot"); mp.bone = None
    sys.modules["matplotlib.pyplot"] = mp
    from common.camera import project_to_2d
    from common.loss import mpjpe_diffusion_reproj, mpjpe_diffusion_all_min
    traj, cam = synthetic_camera(B, F)
    gt = 0.4 * torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(5)); gt[:, :, 0] = 0
    pred = ref.clone(); pred[:, :, :, :, 0] = 0
    b, t_, h, f, j, c = pred.shape
    tr_all = traj.unsqueeze(1).unsqueeze(1).repeat(1, t_, h, 1, 1, 1)
    rep = project_to_2d((pred + tr_all).reshape(b * t_ * h * f, j, c), cam[:1].repeat(b * t_ * h * f, 1)).reshape(b, t_, h, f, j, 2)
    e_ref = mpjpe_diffusion_reproj(pred, gt, rep, x2d)
    p_ref = mpjpe_diffusion_all_min(pred, gt, mean_pos=True)
    jagg, idx, pagg, _ = orc.jpma(ref, traj, cam, x2d)
    e_mine = torch.norm(jagg - gt[:, None], dim=-1).permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    p_mine = torch.norm(pagg - gt[:, None], dim=-1).permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    print("J-Agg per-step error ref", e_ref.tolist(), "oracle", e_mine.tolist())
    print("P-Agg per-step error ref", p_ref.tolist(), "oracle", p_mine.tolist())
    ok &= torch.allclose(e_ref, e_mine, atol=1e-6) and torch.allclose(p_ref, p_mine, atol=1e-6)
    # all four logged errors (main.py:715-718) through the oracle's jpma_errors
    from common.loss import mpjpe_diffusion
    errs = orc.jpma_errors(ref, gt, traj, cam, x2d)
    refs = {"J-Best": mpjpe_diffusion_all_min(pred, gt), "P-Best": mpjpe_diffusion(pred, gt), "P-Agg": p_ref, "J-Agg": e_ref}
    for k, v in refs.items():
        same = torch.allclose(v, errs[k], atol=1e-6)
        print(f"{k:7s} ref {v.tolist()} oracle {errs[k].tolist()} {'ok' if same else 'MISMATCH'}")
        ok &= same
    # Protocol 2 (Procrustes) — the reference goes through numpy float32 SVD; .cuda() calls are routed to CPU here
    import common.loss as ref_loss
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        p2_refs = {"J-Best": ref_loss.p_mpjpe_diffusion_all_min(pred, gt), "P-Best": ref_loss.p_mpjpe_diffusion(pred, gt),
                   "P-Agg": ref_loss.p_mpjpe_diffusion_all_min(pred, gt, mean_pos=True),
                   "J-Agg": ref_loss.p_mpjpe_diffusion_reproj(pred, gt, rep, x2d)}
    finally:
        torch.Tensor.cuda = _cuda
    p2 = orc.p_jpma_errors(ref, gt, idx)
    for k, v in p2_refs.items():
        v = torch.as_tensor(v)
        same = torch.allclose(v.double(), p2[k].double(), atol=2e-6)
        print(f"P2 {k:7s} ref {v.tolist()} oracle {p2[k].tolist()} {'ok' if same else 'MISMATCH'}")
        ok &= same
    # training-mode forward: is_train layout + stochastic depth with injected factors (common/mixste.py:100,114-115,
    # 215-225); q_sample / prepare_diffusion_concat against the reference's (common/diffusionpose.py:260-267,290-306)
    from d3dp_b200.mixste import MixSTE2
    holder = MixSTE2(num_frame=F, embed_dim_ratio=512, depth=8, mlp_ratio=2., drop_path_rate=0.1)
    torch.manual_seed(3)
    masks = holder.draw_drop_masks(B, "cpu")
    assert any((m == 0).any() for m in masks), "no branch dropped: pick another seed"
    x_tr = x_t[:, 0]
    ref_tr = rh.run_reference_train_forward(F, sd, x2d, x_tr, t, masks)
    with torch.no_grad():
        mine_tr = orc.denoiser(sd, x2d, x_tr[:, None], t, drop_masks=masks)[:, 0]
    m, mx = orc.mpjpe_distance(mine_tr, ref_tr)
    print(f"train forward + DropPath  mean {m:.3e} max {mx:.3e}")
    ok &= mx < 1e-5
    tt = torch.tensor([17, 803])
    nz = torch.randn(B, F, 17, 3, generator=torch.Generator().manual_seed(8))
    ref_q = model.q_sample(gt, tt, nz)
    same = torch.equal(ref_q, orc.q_sample(gt, tt, nz))
    print(f"q_sample bit-equal: {same}")
    ok &= same
    print("ORACLE PINNED" if ok else "ORACLE MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
