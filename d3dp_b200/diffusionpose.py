"""Drop-in `D3DP` (reference: common/diffusionpose.py:55-320): same constructor, same registered float64 schedule
buffers, same `state_dict()` keys, same `forward / ddim_sample / ddim_sample_flip / q_sample` surface — but the whole
K-step DDIM loop (denoiser passes, flip test-time augmentation, x0 -> eps, alpha-beta update, stacking of the
per-step predictions) is one call into the sm_100a library (include/d3dp_b200.h: d3dp_ddim_sample).

Host code is PyTorch only for parameters, buffers, memory and the stream.  There is no ATen fallback: on a machine
without the built library or without a Blackwell GPU every sampling call raises.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .mixste import MixSTE2

__all__ = ["D3DP"]


def extract(a, t, x_shape):
    """common/diffusionpose.py:35-39."""
    out = a.gather(-1, t)
    return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def cosine_beta_schedule(timesteps, s=0.008):
    """common/diffusionpose.py:42-52 — float64 cosine schedule."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    acp = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    acp = acp / acp[0]
    return torch.clip(1 - (acp[1:] / acp[:-1]), 0, 0.999)


class D3DP(nn.Module):
    OUTPUT_SCALE = 1.0  # common/diffusionpose_3dhp.py multiplies every returned pose by 1000 (see diffusionpose_3dhp)

    def __init__(self, args, joints_left, joints_right, is_train=True, num_proposals=1, sampling_timesteps=1):
        super().__init__()
        self.frames = args.number_of_frames
        self.num_proposals = num_proposals
        self.flip = args.test_time_augmentation
        self.joints_left = list(joints_left)
        self.joints_right = list(joints_right)
        self.is_train = is_train

        betas = cosine_beta_schedule(args.timestep)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else self.num_timesteps
        assert self.sampling_timesteps <= self.num_timesteps
        self.is_ddim_sampling = self.sampling_timesteps < self.num_timesteps
        self.ddim_sampling_eta = 1.
        self.objective = 'pred_x0'
        self.scale = args.scale

        # the 12 float64 buffers of the reference, in its registration order (common/diffusionpose.py:92-117)
        self.register_buffer('betas', betas)
        self.register_buffer('alphas_cumprod', alphas_cumprod)
        self.register_buffer('alphas_cumprod_prev', alphas_cumprod_prev)
        self.register_buffer('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        self.register_buffer('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        self.register_buffer('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        self.register_buffer('posterior_variance', posterior_variance)
        self.register_buffer('posterior_log_variance_clipped', torch.log(posterior_variance.clamp(min=1e-20)))
        self.register_buffer('posterior_mean_coef1', betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        self.register_buffer('posterior_mean_coef2',
                             (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))

        self.pose_estimator = MixSTE2(
            num_frame=self.frames, num_joints=17, in_chans=2, embed_dim_ratio=args.cs, depth=args.dep, num_heads=8,
            mlp_ratio=2., qkv_bias=True, qk_scale=None, drop_path_rate=0.1 if is_train else 0, is_train=is_train,
            joints_left=self.joints_left, joints_right=self.joints_right, scale=args.scale,
            output_scale=self.OUTPUT_SCALE)
        self._sched_fp = {}
        self._sched = {"master": self, "epoch": 0}  # shared with nn.DataParallel replicas (shallow copies)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._bump_schedule())

    def _bump_schedule(self):
        self._sched["epoch"] += 1

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if "_sched" in self.__dict__:
            self._bump_schedule()
        return out

    # ------------------------------------------------------------------ engine
    def _engine(self):
        eng = self.pose_estimator.engine()
        names = ("alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod")
        # keyed on the module that owns the buffers (DataParallel replicas get fresh broadcast copies every forward):
        # its epoch (load_state_dict / .to() / .cuda()) and the buffers' version counters (in-place edits)
        master = self._sched["master"]
        fp = (self._sched["epoch"],) + tuple(getattr(master, n)._version for n in names)
        if self._sched_fp.get(id(eng)) != fp:
            eng.set_schedule(*(getattr(self, n) for n in names))  # 5 x 1000 doubles, once per change
            self._sched_fp[id(eng)] = fp
        return eng

    def _time_list(self):
        # verbatim semantics of common/diffusionpose.py:221-222 (float32 linspace, .int() truncation)
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        return list(reversed(times.int().tolist()))

    def _sample(self, inputs_2d, input_2d_flip, noise_init, noise_steps, seed, h_offset, H_total):
        eng = self._engine()
        B, H, K = inputs_2d.shape[0], self.num_proposals, self.sampling_timesteps
        shape = (B, H, self.frames, 17, 3)
        if seed is None:  # reference behaviour: draws come from torch's CUDA generator (torch.randn / randn_like)
            if noise_init is None:
                noise_init = torch.randn(shape, device=eng.device)
            if noise_steps is None:
                noise_steps = torch.randn((max(K - 1, 0),) + shape, device=eng.device)
        return eng.ddim_sample(inputs_2d, input_2d_flip, H, K, noise_init=noise_init, noise_steps=noise_steps,
                               seed=seed or 0, h_offset=h_offset, H_total=H_total, timesteps=self._time_list())

    # ------------------------------------------------------------------ reference surface
    def predict_noise_from_start(self, x_t, t, x0):
        """common/diffusionpose.py:129-133 (host-side helper, float64 like the reference)."""
        return ((extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) /
                extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    @torch.no_grad()
    def ddim_sample(self, inputs_2d, inputs_3d, clip_denoised=True, do_postprocess=True, *, noise_init=None,
                    noise_steps=None, seed=None, h_offset=0, H_total=None):
        """common/diffusionpose.py:172-212: no-flip sampler; returns the list of K x_start tensors [B,H,F,17,3].
        (The reference crashes for K >= 2 on a float64 promotion; here `img` stays float32 as in the flip path.)"""
        preds = self._sample(inputs_2d, None, noise_init, noise_steps, seed, h_offset, H_total)
        return [preds[:, k] for k in range(preds.shape[1])]

    @torch.no_grad()
    def ddim_sample_flip(self, inputs_2d, inputs_3d, clip_denoised=True, do_postprocess=True, input_2d_flip=None, *,
                         noise_init=None, noise_steps=None, seed=None, h_offset=0, H_total=None):
        """common/diffusionpose.py:215-256: returns torch.stack(preds_all, dim=1) = [B,K,H,F,17,3].
        Keyword-only extras: injected noise (parity runs), or `seed` for in-kernel Philox noise addressed by the
        global hypothesis index h_offset+h of H_total (multi-GPU hypothesis sharding)."""
        if input_2d_flip is None:
            raise ValueError("ddim_sample_flip needs input_2d_flip (as the reference does)")
        return self._sample(inputs_2d, input_2d_flip, noise_init, noise_steps, seed, h_offset, H_total)

    @torch.no_grad()
    def q_sample(self, x_start, t, noise=None):
        """common/diffusionpose.py:260-267 (returned as float32; the reference's float64 promotion is rounded once)."""
        if noise is None:
            noise = torch.randn_like(x_start)
        return self._engine().q_sample(x_start, noise, t, clamp=False)

    @torch.no_grad()
    def prepare_targets(self, targets):
        """common/diffusionpose.py:290-320: per-sample t ~ U[0,T), noising, clamp(+-1.1 scale)/scale."""
        eng = self._engine()
        B = targets.shape[0]
        t = torch.randint(0, self.num_timesteps, (B,), device=eng.device).long()
        noise = torch.randn(B, self.frames, 17, 3, device=eng.device)
        x = eng.q_sample(targets, noise, t, clamp=True)
        return x, noise, t[:, None]

    def forward(self, input_2d, input_3d, input_2d_flip=None):
        """common/diffusionpose.py:269-287."""
        if not self.is_train:
            if self.flip:
                return self.ddim_sample_flip(input_2d, input_3d, input_2d_flip=input_2d_flip)
            return self.ddim_sample(input_2d, input_3d)
        x_poses, noises, t = self.prepare_targets(input_3d / self.OUTPUT_SCALE if self.OUTPUT_SCALE != 1.0 else input_3d)
        pred = self.pose_estimator(input_2d, x_poses.float(), t.squeeze(-1))
        return pred * self.OUTPUT_SCALE if self.OUTPUT_SCALE != 1.0 else pred
