"""Evaluation metrics of the reference's JPMA report, computed from the outputs of the fused kernels
(`Engine.jpma_gt` -> jpma_kernel, `Engine.pmpjpe` -> procrustes_kernel): all per-element work (projection, 2-D and 3-D
errors, argmin, gathers, means over hypotheses, per-pose rigid alignment) runs in those kernels; only the final
means / minima over small tensors are torch reductions on the device.  Nothing goes through numpy or the host.

  Protocol 1 (main.py:715-718):  mpjpe_diffusion_all_min, mpjpe_diffusion, mpjpe_diffusion_all_min(mean_pos=True),
                                 mpjpe_diffusion_reproj                                  (common/loss.py:22-107)
  Protocol 2 (main.py:726-729):  p_mpjpe_diffusion_all_min, p_mpjpe_diffusion, ...(mean_pos=True),
                                 p_mpjpe_diffusion_reproj                                 (common/loss.py:190-395)
  3DHP (main_3dhp.py:851-852):   mpjpe_diffusion_3dhp with the valid-frame mask           (common/loss.py:109-145)
"""
import torch


def _per_step(e, K):
    """[B,K,...] -> [K] mean over everything but the step axis."""
    return e.transpose(0, 1).reshape(K, -1).mean(-1)


def _best_over_h(e, K, H):
    """e [B,K,H,F,17] -> P-Best error [K] and its hypothesis index [K] (mean over b,f,n per h, then min over h)."""
    per_h = e.permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(-1)
    m = per_h.min(dim=1)
    return m.values, m.indices


def jpma_metrics(engine, preds, gt, traj, cam, x2d, root_joint=0, linear=False, protocol2=False):
    """Returns {"J-Best", "P-Best", "P-Agg", "J-Agg"}: tensors [K] (error per DDIM step, in the units of `preds`),
    plus the kernel outputs (jagg_pose, jagg_idx, pagg_pose, jbest_pose, e3d, e2d_min) and the P-Best index.
    `preds` [B,K,H,F,17,3] as returned by the sampler (the root joint is zeroed inside the kernels, main.py:700);
    `gt` [B,F,17,3] with its root joint already zeroed (main.py:683).  With protocol2=True the dict also holds
    "P2-J-Best", "P2-P-Best", "P2-P-Agg", "P2-J-Agg" (main.py:726-729)."""
    out = engine.jpma_gt(preds, traj, cam, x2d, gt, root_joint=root_joint, linear=linear)
    K, H = preds.shape[1], preds.shape[2]
    gt = gt.to(out["e3d"].device, torch.float32)
    e3d = out["e3d"]                                                      # [B,K,H,F,17]
    sel = out["jagg_idx"].long().unsqueeze(2)
    res = dict(out)
    res["J-Best"] = _per_step(e3d.min(dim=2).values, K)
    res["P-Best"], res["pbest_idx"] = _best_over_h(e3d, K, H)
    res["P-Agg"] = _per_step(torch.norm(out["pagg_pose"] - gt[:, None], dim=-1), K)
    res["J-Agg"] = _per_step(torch.gather(e3d, 2, sel), K)
    if protocol2:
        pe = engine.pmpjpe(preds, gt, root_joint=root_joint)              # [B,K,H,F,17]
        pe_mean = engine.pmpjpe(out["pagg_pose"], gt, root_joint=-1)      # [B,K,F,17]  (root already zero)
        res["pe3d"], res["pe3d_mean"] = pe, pe_mean
        res["P2-J-Best"] = _per_step(pe.min(dim=2).values, K)
        res["P2-P-Best"], _ = _best_over_h(pe, K, H)
        res["P2-P-Agg"] = _per_step(pe_mean, K)
        res["P2-J-Agg"] = _per_step(torch.gather(pe, 2, sel), K)
    return res


def pbest_pose(preds, pbest_idx, root_joint=0):
    """P-Best pose export (main_3dhp.py:785-795): per step, the single hypothesis with the lowest mean error over the
    whole batch.  preds [B,K,H,F,17,3], pbest_idx [K] -> [B,K,F,17,3]."""
    B, K, H, F = preds.shape[:4]
    idx = pbest_idx.to(preds.device).reshape(1, K, 1, 1, 1, 1).expand(B, K, 1, F, 17, 3)
    pose = torch.gather(preds, 2, idx).squeeze(2).clone()
    pose[..., root_joint, :] = 0
    return pose


def valid_frame_metrics(e3d, pagg_pose, gt, valid):
    """3DHP errors restricted to valid frames (mpjpe_diffusion_3dhp, common/loss.py:109-145): e3d [B,K,H,F,17],
    pagg_pose [B,K,F,17,3], gt [B,F,17,3], valid [B,F] or [B,F,1] bool.  Returns (P-Best [K], P-Agg [K]); with no
    valid frame at all the result is NaN, as the reference's empty mean is."""
    B, K, H, F = e3d.shape[:4]
    valid = valid.reshape(B, F).to(e3d.device, torch.bool)
    w = valid[:, None, None, :, None].to(e3d.dtype)
    n = valid.sum() * 17
    per_h = (e3d * w).permute(1, 2, 0, 3, 4).reshape(K, H, -1).sum(-1) / n
    e_mean = torch.norm(pagg_pose - gt[:, None].to(pagg_pose.device), dim=-1) * w[:, :, 0]
    return per_h.min(dim=1).values, e_mean.transpose(0, 1).reshape(K, -1).sum(-1) / n
