"""Evaluation metrics of the reference's JPMA report (main.py:715-718 → common/loss.py:22-107), computed from the
outputs of the fused JPMA kernel (`Engine.jpma_gt`): the per-element work (projection, 2-D and 3-D errors, argmin,
gather, mean) is done in `jpma_kernel`; only the final means / minima over small tensors are torch reductions."""
import torch


def jpma_metrics(engine, preds, gt, traj, cam, x2d, root_joint=0, linear=False):
    """Returns {"J-Best", "P-Best", "P-Agg", "J-Agg"}: tensors [K] (error per DDIM step, model units), matching
    mpjpe_diffusion_all_min, mpjpe_diffusion, mpjpe_diffusion_all_min(mean_pos=True), mpjpe_diffusion_reproj.
    `preds` [B,K,H,F,17,3] as returned by the sampler (the root joint is zeroed inside the kernel, main.py:700);
    `gt` [B,F,17,3] with its root joint already zeroed (main.py:683)."""
    out = engine.jpma_gt(preds, traj, cam, x2d, gt, root_joint=root_joint, linear=linear)
    K = preds.shape[1]
    gt = gt.to(out["e3d"].device, torch.float32)
    e3d = out["e3d"]                                                      # [B,K,H,F,17]
    j_best = e3d.min(dim=2).values.permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    p_best = e3d.permute(1, 2, 0, 3, 4).reshape(K, e3d.shape[2], -1).mean(-1).min(dim=1).values
    p_agg = torch.norm(out["pagg_pose"] - gt[:, None], dim=-1).permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    j_agg = torch.norm(out["jagg_pose"] - gt[:, None], dim=-1).permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    return {"J-Best": j_best, "P-Best": p_best, "P-Agg": p_agg, "J-Agg": j_agg, **out}
