"""CPU oracle for the D3DP diffusion-sampling hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain float32 (float64 where the reference is float64) restatement, on the CPU, of what the reference computes on
this path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs (CPU baseline, `--impl
reference`, and the same-GPU eager-PyTorch baseline, which runs these same functions on CUDA tensors) may import it;
nothing under `d3dp_b200/` does.

The reference's arithmetic for this path *is* PyTorch ATen (addmm / bmm / softmax / layer_norm / gelu; SURVEY §8c:
torch unpinned in the reference, 2.11.0 here), so the restatement uses the same ATen ops on CPU tensors, but in one
canonical [B, H, F, 17, C] layout without the reference's rearrange/permute copies, and as pure functions of a
state_dict.  It is pinned against the reference itself: `oracle/validate_against_reference.py` runs the unmodified
`/root/reference` D3DP (with the timm stub and noise-injection proxy of SURVEY Appendix C) on the same weights,
inputs and noise and requires agreement to float32 re-association noise; `tests/golden/` holds outputs produced by
that same reference run, so the pin travels to machines without /root/reference.

Each function cites the reference lines it follows.
"""
import math

import numpy as np
import torch
import torch.nn.functional as Fn

J = 17


# ----------------------------------------------------------------------------------------------- schedule
def cosine_beta_schedule(timesteps, s=0.008):
    """common/diffusionpose.py:42-52 (float64)."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


def schedule_buffers(timesteps=1000):
    """The float64 buffers D3DP registers (common/diffusionpose.py:75-117)."""
    betas = cosine_beta_schedule(timesteps)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = Fn.pad(ac[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(post_var.clamp(min=1e-20)),
        "posterior_mean_coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
    }


def time_list(timesteps, K):
    """common/diffusionpose.py:221-222: reversed(linspace(-1, T-1, K+1).int())."""
    times = torch.linspace(-1, timesteps - 1, steps=K + 1)
    return list(reversed(times.int().tolist()))


# ----------------------------------------------------------------------------------------------- denoiser
def _block(x, sd, pre, drop=None):
    """Block.forward + Attention.forward + Mlp.forward on x[..., N, 512] (common/mixste.py:63-82,113-115,37-43).
    `drop` = (attention factor, mlp factor), each broadcastable to x[..., :1, :1]: timm's DropPath multiplies the
    branch output by bernoulli(keep)/keep per sample of the block input's first axis (training only)."""
    C = x.shape[-1]
    h = Fn.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
    qkv = Fn.linear(h, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
    lead, N = x.shape[:-2], x.shape[-2]
    qkv = qkv.reshape(*lead, N, 3, 8, C // 8)
    q, k, v = (qkv[..., i, :, :].transpose(-2, -3) for i in range(3))  # [..., 8, N, 64]
    att = (q @ k.transpose(-2, -1)) * ((C // 8) ** -0.5)
    att = att.softmax(dim=-1)
    o = (att @ v).transpose(-2, -3).reshape(*lead, N, C)
    a = Fn.linear(o, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    x = x + (a if drop is None else a * drop[0])
    h = Fn.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)
    h = Fn.gelu(Fn.linear(h, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]))
    m = Fn.linear(h, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    return x + (m if drop is None else m * drop[1])


def time_embedding(sd, t):
    """SinusoidalPositionEmbeddings + time_mlp (common/mixste.py:127-139,179-184). t: int64 [B] -> [B,512]."""
    half = 256
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=t.device) * -e)
    e = t[:, None] * e[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = Fn.gelu(Fn.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return Fn.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def denoiser(sd, x2d, x_t, t, depth=8, taps=None, drop_masks=None):
    """MixSTE2.forward, eval branch (common/mixste.py:226-298): x2d [B,F,17,2], x_t [B,H,F,17,3], t [B] int64
    -> [B,H,F,17,3].  `taps`, if a dict, receives intermediate activations for layer-by-layer checks.
    `drop_masks` (training-mode forward): depth x 4 DropPath factor tensors in execution order — STEblocks[d]
    attention [B*H,F], STEblocks[d] mlp [B*H,F], TTEblocks[d] attention [B*H,17], TTEblocks[d] mlp [B*H,17]."""
    B, H, F = x_t.shape[0], x_t.shape[1], x_t.shape[2]
    C = 512
    u = torch.cat((x2d[:, None].expand(B, H, F, J, 2), x_t), dim=-1)                       # :227-228
    x = Fn.linear(u, sd["Spatial_patch_to_embedding.weight"], sd["Spatial_patch_to_embedding.bias"])
    x = x + sd["Spatial_pos_embed"].reshape(1, 1, 1, J, C)                                   # :232
    x = x + time_embedding(sd, t)[:, None, None, None, :]                                   # :233-235
    if taps is not None:
        taps["embed"] = x.clone()
    ln_s = (sd["Spatial_norm.weight"], sd["Spatial_norm.bias"])
    ln_t = (sd["Temporal_norm.weight"], sd["Temporal_norm.bias"])
    def drops(d, which, n):
        if drop_masks is None:
            return None
        return tuple(drop_masks[4 * d + 2 * which + i].reshape(B, H, n, 1, 1).to(x.dtype) for i in range(2))
    for d in range(depth):
        x = _block(x, sd, f"STEblocks.{d}.", drops(d, 0, F))                                # over the 17 joints
        x = Fn.layer_norm(x, (C,), ln_s[0], ln_s[1], 1e-6)                                  # :243,269
        if d == 0:
            x = x + sd["Temporal_pos_embed"].reshape(1, 1, F, 1, C)                         # :250
        if taps is not None:
            taps[f"S{d}"] = x.clone()
        xt = x.transpose(2, 3)                                                              # [B,H,17,F,C]
        xt = _block(xt, sd, f"TTEblocks.{d}.", drops(d, 1, J))                              # over the F frames
        xt = Fn.layer_norm(xt, (C,), ln_t[0], ln_t[1], 1e-6)                                # :257,273
        x = xt.transpose(2, 3)
        if taps is not None:
            taps[f"T{d}"] = x.clone()
    x = Fn.layer_norm(x, (C,), sd["head.0.weight"], sd["head.0.bias"], 1e-5)                # :207-210
    return Fn.linear(x, sd["head.1.weight"], sd["head.1.bias"])


# ----------------------------------------------------------------------------------------------- sampler
def flip_pose(x, joints_left, joints_right):
    """common/diffusionpose.py:150-153 (an involution; also used to un-flip, :158-160)."""
    y = x.clone()
    y[..., 0] *= -1
    y[..., joints_left + joints_right, :] = y[..., joints_right + joints_left, :]
    return y


def ddim_sample(sd, x2d, x2d_flip, H, K, noise_init, noise_steps, joints_left, joints_right, scale=1.0,
                timesteps=1000, depth=8, buffers=None):
    """D3DP.ddim_sample_flip (common/diffusionpose.py:215-256) when x2d_flip is given, D3DP.ddim_sample (:172-212,
    with the float32 result the reference would give if it did not crash on its float64 promotion) otherwise.
    noise_init [B,H,F,17,3], noise_steps [K-1,B,H,F,17,3] are the injected randn draws.  Returns [B,K,H,F,17,3]."""
    bufs = buffers or schedule_buffers(timesteps)
    ac = bufs["alphas_cumprod"]
    B = x2d.shape[0]
    times = time_list(timesteps, K)
    img = noise_init.clone()
    preds = []
    for k, (t, t_next) in enumerate(zip(times[:-1], times[1:])):
        tc = torch.full((B,), t, dtype=torch.long, device=x2d.device)
        x_t = torch.clamp(img, min=-1.1 * scale, max=1.1 * scale) / scale                   # :148-149
        pred = denoiser(sd, x2d, x_t, tc, depth)
        if x2d_flip is not None:
            pf = denoiser(sd, x2d_flip, flip_pose(x_t, joints_left, joints_right), tc, depth)
            pred = (pred + flip_pose(pf, joints_left, joints_right)) / 2                     # :158-161
        x0 = torch.clamp(pred * scale, min=-1.1 * scale, max=1.1 * scale)                   # :163-165
        # predict_noise_from_start with float64 buffers, then .float()  (:129-133,166-167)
        eps = ((bufs["sqrt_recip_alphas_cumprod"][t] * img.double() - x0.double()) /
               bufs["sqrt_recipm1_alphas_cumprod"][t]).float()
        preds.append(x0)
        if t_next < 0:
            img = x0
            continue
        a, an = ac[t], ac[t_next]
        sigma = 1.0 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()                             # :247
        c = (1 - an - sigma ** 2).sqrt()                                                    # :248
        img = x0 * an.sqrt() + c * eps + sigma * noise_steps[k]                              # :252-254 (0-dim f64 scalars)
    return torch.stack(preds, dim=1)


def q_sample(x0, t, noise, buffers=None, timesteps=1000):
    """common/diffusionpose.py:260-267 (float64 result like the reference's promoted expression)."""
    bufs = buffers or schedule_buffers(timesteps)
    shape = (x0.shape[0],) + (1,) * (x0.dim() - 1)
    return bufs["sqrt_alphas_cumprod"][t].reshape(shape) * x0 + \
        bufs["sqrt_one_minus_alphas_cumprod"][t].reshape(shape) * noise


def prepare_diffusion(pose, t, noise, scale=1.0, buffers=None):
    """common/diffusionpose.py:290-306 for a batch: scale, q_sample, clamp, unscale (then .float(), :281)."""
    x = q_sample(pose * scale, t, noise, buffers)
    x = torch.clamp(x, min=-1.1 * scale, max=1.1 * scale) / scale
    return x.float()


# ----------------------------------------------------------------------------------------------- JPMA
def project_to_2d(X, cam):
    """common/camera.py:44-60. X [N,*,3], cam [N,9]."""
    while cam.dim() < X.dim():
        cam = cam.unsqueeze(1)
    f, c, k, p = cam[..., :2], cam[..., 2:4], cam[..., 4:7], cam[..., 7:]
    XX = torch.clamp(X[..., :2] / X[..., 2:], min=-1, max=1)
    r2 = torch.sum(XX[..., :2] ** 2, dim=-1, keepdim=True)
    radial = 1 + torch.sum(k * torch.cat((r2, r2 ** 2, r2 ** 3), dim=-1), dim=-1, keepdim=True)
    tan = torch.sum(p * XX, dim=-1, keepdim=True)
    return f * (XX * (radial + tan) + p * r2) + c


def project_to_2d_linear(X, cam):
    """common/camera.py:62-80."""
    while cam.dim() < X.dim():
        cam = cam.unsqueeze(1)
    f, c = cam[..., :2], cam[..., 2:4]
    XX = torch.clamp(X[..., :2] / X[..., 2:], min=-1, max=1)
    return f * XX + c


def jpma(preds, traj, cam, x2d, root_joint=0, linear=False):
    """main.py:700-712 + common/loss.py:54-76 + main_3dhp.py:782,801-835.
    preds [B,K,H,F,17,3], traj [B,F,1,3], cam [B,9], x2d [B,F,17,2]
    -> J-Agg pose [B,K,F,17,3], J-Agg index [B,K,F,17], P-Agg pose [B,K,F,17,3], min 2-D error [B,K,F,17]."""
    B, K, H, F = preds.shape[:4]
    P = preds.clone()
    P[:, :, :, :, root_joint] = 0
    X = P + traj.reshape(B, 1, 1, F, 1, 3)
    proj = project_to_2d_linear if linear else project_to_2d
    uv = proj(X.reshape(B, K * H * F, J, 3), cam).reshape(B, K, H, F, J, 2)
    e2d = torch.norm(uv - x2d.reshape(B, 1, 1, F, J, 2), dim=-1)                             # [B,K,H,F,17]
    mn = torch.min(e2d, dim=2, keepdim=True)
    idx = mn.indices                                                                        # [B,K,1,F,17]
    jagg = torch.gather(P, 2, idx.unsqueeze(-1).expand(B, K, 1, F, J, 3)).squeeze(2)
    pagg = torch.mean(P, dim=2)
    return jagg, idx.squeeze(2), pagg, mn.values.squeeze(2)


def jpma_errors(preds, gt, traj, cam, x2d, root_joint=0, linear=False):
    """The four per-step errors main.py:715-718 logs — J-Best (common/loss.py:22-40 mpjpe_diffusion_all_min), P-Best
    (:78-94 mpjpe_diffusion), P-Agg (:42-52 mean_pos branch), J-Agg (:54-76 mpjpe_diffusion_reproj) — as [K] tensors.
    preds [B,K,H,F,17,3] (root joint zeroed here like main.py:700), gt [B,F,17,3] with the root joint zeroed."""
    B, K, H, F = preds.shape[:4]
    P = preds.clone()
    P[:, :, :, :, root_joint] = 0
    e3d = torch.norm(P - gt.reshape(B, 1, 1, F, J, 3), dim=-1)                                   # [B,K,H,F,17]
    j_best = e3d.permute(1, 2, 0, 3, 4).min(dim=1).values.reshape(K, -1).mean(-1)
    p_best = e3d.permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(-1).min(dim=1).values
    jagg, idx, pagg, _ = jpma(preds, traj, cam, x2d, root_joint, linear)
    p_agg = torch.norm(pagg - gt[:, None], dim=-1).permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    sel = torch.gather(e3d, 2, idx.unsqueeze(2))                                                # error of the selected h
    j_agg = sel.permute(1, 2, 0, 3, 4).reshape(K, -1).mean(-1)
    jbest_idx = e3d.min(dim=2, keepdim=True).indices
    jbest_pose = torch.gather(P, 2, jbest_idx.unsqueeze(-1).expand(B, K, 1, F, J, 3)).squeeze(2)
    return {"J-Best": j_best, "P-Best": p_best, "P-Agg": p_agg, "J-Agg": j_agg, "e3d": e3d, "jbest_pose": jbest_pose}


def procrustes_errors(pred, gt):
    """Per-joint distance after the rigid alignment of common/loss.py:208-238 (shared by the three p_mpjpe_* variants):
    pred [..., 17, 3] poses, gt broadcastable to it.  float64 torch.linalg.svd instead of numpy's float32 LAPACK."""
    Y = pred.double().reshape(-1, J, 3)
    X = gt.double().expand_as(pred).reshape(-1, J, 3)
    muX, muY = X.mean(1, keepdim=True), Y.mean(1, keepdim=True)
    X0, Y0 = X - muX, Y - muY
    nX = X0.pow(2).sum((1, 2), keepdim=True).sqrt()
    nY = Y0.pow(2).sum((1, 2), keepdim=True).sqrt()
    M = (X0 / nX).transpose(1, 2) @ (Y0 / nY)
    U, s, Vt = torch.linalg.svd(M)
    V = Vt.transpose(1, 2).clone()
    d = torch.sign(torch.linalg.det(V @ U.transpose(1, 2)))
    V[:, :, -1] *= d[:, None]
    s = s.clone()
    s[:, -1] *= d
    R = V @ U.transpose(1, 2)
    a = s.sum(1)[:, None, None] * nX / nY
    t = muX - a * (muY @ R)
    aligned = a * (Y @ R) + t
    return torch.norm(aligned - X, dim=-1).reshape(pred.shape[:-1]).float()


def p_jpma_errors(preds, gt, jagg_idx, root_joint=0):
    """Protocol-2 versions of the four logged errors (main.py:726-729): preds [B,K,H,F,17,3], gt [B,F,17,3] (root
    zeroed), jagg_idx [B,K,F,17] = the reprojection argmin of jpma().  Returns [K] tensors + the per-pose errors."""
    B, K, H, F = preds.shape[:4]
    P = preds.clone()
    P[:, :, :, :, root_joint] = 0
    pe = procrustes_errors(P, gt.reshape(B, 1, 1, F, J, 3))                                      # [B,K,H,F,17]
    pe_mean = procrustes_errors(P.mean(dim=2), gt.reshape(B, 1, F, J, 3))                        # [B,K,F,17]
    j_best = pe.permute(1, 2, 0, 3, 4).min(dim=1).values.reshape(K, -1).mean(-1)
    p_best = pe.permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(-1).min(dim=1).values
    p_agg = pe_mean.permute(1, 0, 2, 3).reshape(K, -1).mean(-1)
    j_agg = torch.gather(pe, 2, jagg_idx.long().unsqueeze(2)).permute(1, 2, 0, 3, 4).reshape(K, -1).mean(-1)
    return {"J-Best": j_best, "P-Best": p_best, "P-Agg": p_agg, "J-Agg": j_agg, "pe3d": pe, "pe3d_mean": pe_mean}


def eval_data_prepare(receptive_field, inputs_2d):
    """main.py:267-299 for one sequence [N,17,C]: ceil(N/F) clips, the last one = the last F frames; replicate-pad
    sequences shorter than F."""
    F = receptive_field
    x = inputs_2d
    n = x.shape[0]
    out_num = n // F + (1 if n % F else 0)
    out = torch.empty(max(out_num, 1), F, x.shape[1], x.shape[2])
    for i in range(out_num - 1):
        out[i] = x[i * F:(i + 1) * F]
    if n < F:
        x = torch.cat([x, x[-1:].repeat(F - n, 1, 1)], dim=0)
    out[-1] = x[-F:]
    return out


def pose_post_process(pose_pred, n_frames, receptive_field):
    """main_3dhp.py:327-332 (+ :717-718 allocation): per-clip poses [n_clips, K, F, 17, 3] of one sequence are written
    clip by clip into a zero [K, N, 17, 3] buffer, the last clip then overwrites the LAST F frames, and the buffer is
    transposed to the MATLAB layout [3, 17, N, K]."""
    F = receptive_field
    pose_pred = np.asarray(pose_pred)
    buf = np.zeros((pose_pred.shape[1], n_frames, J, 3))
    for ii in range(pose_pred.shape[0] - 1):
        buf[:, ii * F:(ii + 1) * F] = pose_pred[ii]
    buf[:, -F:] = pose_pred[-1][:, -min(F, n_frames):] if n_frames < F else pose_pred[-1]
    return buf.transpose(3, 2, 1, 0)


def mpjpe_3dhp_valid(predicted, target, valid_frame, mean_pos=False):
    """common/loss.py:109-145 mpjpe_diffusion_3dhp: P-Best (mean_pos=False) / P-Agg (mean_pos=True) error per step
    over the VALID frames only.  predicted [B,K,H,F,17,3], target [B,F,17,3], valid_frame [B,F,1] bool -> [K]."""
    valid = valid_frame.squeeze(2)
    pv = predicted.permute(0, 3, 1, 2, 4, 5)[valid]                       # [n, K, H, 17, 3]
    tv = target[valid]                                                    # [n, 17, 3]
    K, H = pv.shape[1], pv.shape[2]
    if not mean_pos:
        e = torch.norm(pv - tv[:, None, None], dim=-1)                    # [n, K, H, 17]
        e = e.permute(1, 2, 0, 3).reshape(K, H, -1).mean(-1)
        return e.min(dim=1).values
    e = torch.norm(pv.mean(dim=2) - tv[:, None], dim=-1)                  # [n, K, 17]
    return e.permute(1, 0, 2).reshape(K, -1).mean(-1)


def image_coordinates(x, w, h):
    """common/camera.py:14-18."""
    return (x + torch.tensor([1.0, h / w], dtype=x.dtype)) * w / 2


def pbest_pose(preds, gt):
    """main_3dhp.py:785-795: the single hypothesis with the lowest batch-mean error per step, gathered for every
    clip.  preds [B,K,H,F,17,3] (root already zeroed), gt [B,F,17,3] -> pose [B,K,F,17,3], index [K]."""
    B, K, H, F = preds.shape[:4]
    e = torch.norm(preds - gt[:, None, None], dim=-1)
    eh = e.permute(1, 2, 0, 3, 4).reshape(K, H, -1).mean(-1, keepdim=True)
    idx = eh.min(dim=1, keepdim=True).indices                              # [K,1,1]
    g = idx.unsqueeze(0).unsqueeze(-1).unsqueeze(-1).repeat(B, 1, 1, F, J, 3)
    return torch.gather(preds, 2, g).squeeze(2), idx.reshape(K)


def mpjpe_distance(a, b):
    """Parity metric: mean / max over all joints of the per-joint L2 distance."""
    d = torch.norm(a.double() - b.double(), dim=-1)
    return d.mean().item(), d.max().item()


# ----------------------------------------------------------------------------------------------- Philox (integer part)
def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al. 2011) on uint32 numpy arrays: ctr [..,4], key (k0,k1). Bit-exact spec for the
    counter-based noise in d3dp_b200/csrc/elementwise.cuh."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack([x.astype(np.uint32) for x in c], axis=-1)


def philox_normal(seed, draw, elem):
    """float64 restatement of philox_normal(): Box-Muller on words 0,1. elem: uint64 numpy array."""
    elem = np.asarray(elem, dtype=np.uint64)
    ctr = np.stack([(elem & np.uint64(0xFFFFFFFF)).astype(np.uint32), (elem >> np.uint64(32)).astype(np.uint32),
                    np.full(elem.shape, draw, np.uint32), np.zeros(elem.shape, np.uint32)], axis=-1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u1 = ((r[..., 0] >> 8).astype(np.float64) + 0.5) / 16777216.0
    u2 = ((r[..., 1] >> 8).astype(np.float64) + 0.5) / 16777216.0
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
