"""Thin Python owner of a `d3dp_handle` (include/d3dp_b200.h): weight upload, workspace, and one method per C entry
point.  PyTorch is used only for device memory and the CUDA stream; every computation is inside libd3dp_b200.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import D3dpConfig, D3dpError, check, ptr

H36M_JOINTS_LEFT = [4, 5, 6, 11, 12, 13]
H36M_JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]


def flip_permutation(joints_left, joints_right, n=17):
    """perm[j] = source joint of j under x[..., L+R, :] = x[..., R+L, :] (common/diffusionpose.py:152-153)."""
    perm = list(range(n))
    for dst, src in zip(list(joints_left) + list(joints_right), list(joints_right) + list(joints_left)):
        perm[dst] = src
    return perm


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    def __init__(self, frames, joints_left=H36M_JOINTS_LEFT, joints_right=H36M_JOINTS_RIGHT, depth=8, channels=512,
                 scale=1.0, num_timesteps=1000, device=None, output_scale=1.0):
        if not torch.cuda.is_available():
            raise D3dpError("d3dp_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.frames, self.depth, self.scale, self.num_timesteps = frames, depth, float(scale), num_timesteps
        cfg = D3dpConfig()
        cfg.frames, cfg.joints, cfg.channels, cfg.depth = frames, 17, channels, depth
        cfg.heads, cfg.mlp_hidden, cfg.num_timesteps, cfg.scale = 8, 2 * channels, num_timesteps, float(scale)
        for j, s in enumerate(flip_permutation(joints_left, joints_right)):
            cfg.flip_perm[j] = s
        cfg.output_scale = float(output_scale)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.d3dp_create(C.byref(cfg), C.byref(self.handle))
        if rc != 0:
            raise D3dpError(f"d3dp_create failed (code {rc}): needs an sm_100 GPU, C=512, F<=384, depth<=8")
        self._ws = None
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.d3dp_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights / schedule
    def load_pose_estimator_state(self, state):
        """`state`: {name: tensor} with the reference's `pose_estimator.*` keys (prefix stripped)."""
        with torch.cuda.device(self.device):
            for name, t in state.items():
                t = t.detach().to(self.device, torch.float32).contiguous()
                check(self.handle, self.lib.d3dp_set_weight(self.handle, name.encode(), ptr(t), t.numel(), _stream()),
                      f"d3dp_set_weight({name})")
            torch.cuda.current_stream().synchronize()
        missing = self.lib.d3dp_weights_missing(self.handle)
        if missing:
            raise D3dpError(f"{missing} weight tensors missing after load")

    def set_schedule(self, alphas_cumprod, sqrt_recip, sqrt_recipm1, sqrt_ac, sqrt_1mac):
        arrs = [a.detach().to("cpu", torch.float64).contiguous()
                for a in (alphas_cumprod, sqrt_recip, sqrt_recipm1, sqrt_ac, sqrt_1mac)]
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_set_schedule(self.handle, *[ptr(a) for a in arrs], arrs[0].numel(),
                                                          _stream()), "d3dp_set_schedule")

    def alphas_cumprod(self):
        out = torch.empty(self.num_timesteps, dtype=torch.float64)
        check(self.handle, self.lib.d3dp_get_alphas_cumprod(self.handle, ptr(out), out.numel()), "get_alphas_cumprod")
        return out

    # ------------------------------------------------------------------ workspace
    def workspace(self, B, H, flip):
        n = C.c_size_t()
        check(self.handle, self.lib.d3dp_workspace_bytes(self.handle, B, H, int(flip), C.byref(n)), "workspace_bytes")
        if self._ws is None or self._ws.numel() < n.value:
            self._ws = None
            self._ws = torch.empty(n.value, dtype=torch.uint8, device=self.device)
        return self._ws

    def _f32(self, t):
        return t.detach().to(self.device, torch.float32).contiguous()

    @staticmethod
    def _expect(t, shape, what):
        """The C ABI takes raw pointers: a tensor of the wrong shape would be read out of bounds, so refuse it here."""
        if t is not None and tuple(t.shape) != tuple(shape):
            raise D3dpError(f"{what}: expected shape {tuple(shape)}, got {tuple(t.shape)}")

    # ------------------------------------------------------------------ entry points
    def drop_scale_numel(self, n_streams):
        """Elements of the packed DropPath factor buffer d3dp_denoise takes (include/d3dp_b200.h)."""
        return self.depth * 2 * n_streams * (self.frames + 17)

    def denoise(self, x2d, x_t, t, drop_scale=None):
        x2d, x_t = self._f32(x2d), self._f32(x_t)
        t = t.detach().to(self.device, torch.int64).contiguous()
        B, H = x_t.shape[0], x_t.shape[1]
        self._expect(x2d, (B, self.frames, 17, 2), "x_2d")
        self._expect(x_t, (B, H, self.frames, 17, 3), "x_3d")
        self._expect(t, (B,), "t")
        if drop_scale is not None:
            drop_scale = self._f32(drop_scale)
            assert drop_scale.numel() == self.drop_scale_numel(B * H)
        out = torch.empty_like(x_t)
        ws = self.workspace(B, H, False)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_denoise(self.handle, ptr(x2d), ptr(x_t), ptr(t), ptr(drop_scale), ptr(out),
                                                     B, H, ptr(ws), ws.numel(), _stream()), "d3dp_denoise")
        return out

    def ddim_sample(self, x2d, x2d_flip, H, K, noise_init=None, noise_steps=None, seed=0, h_offset=0, H_total=None,
                    timesteps=None):
        x2d = self._f32(x2d)
        x2d_flip = None if x2d_flip is None else self._f32(x2d_flip)
        noise_init = None if noise_init is None else self._f32(noise_init)
        noise_steps = None if noise_steps is None else self._f32(noise_steps)
        B = x2d.shape[0]
        H_total = H if H_total is None else H_total
        self._expect(x2d, (B, self.frames, 17, 2), "inputs_2d")
        self._expect(x2d_flip, (B, self.frames, 17, 2), "input_2d_flip")
        self._expect(noise_init, (B, H, self.frames, 17, 3), "noise_init")
        self._expect(noise_steps, (max(K - 1, 0), B, H, self.frames, 17, 3), "noise_steps")
        preds = torch.empty(B, K, H, self.frames, 17, 3, dtype=torch.float32, device=self.device)
        ws = self.workspace(B, H, x2d_flip is not None)
        ts = None
        if timesteps is not None:
            ts = torch.tensor(list(timesteps), dtype=torch.int32)
            assert ts.numel() == K + 1
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_ddim_sample(
                self.handle, ptr(x2d), ptr(x2d_flip), ptr(noise_init), ptr(noise_steps), int(seed), int(h_offset),
                int(H_total), ptr(ts), ptr(preds), B, H, K, ptr(ws), ws.numel(), _stream()), "d3dp_ddim_sample")
        return preds

    def q_sample(self, x0, noise, t, clamp=False):
        x0, noise = self._f32(x0), self._f32(noise)
        t = t.detach().to(self.device, torch.int64).contiguous()
        B = x0.shape[0]
        out = torch.empty_like(x0)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_q_sample(self.handle, ptr(x0), ptr(noise), ptr(t), ptr(out), B,
                                                      x0.numel() // B, int(clamp), _stream()), "d3dp_q_sample")
        return out

    def jpma(self, preds, traj, cam, x2d, root_joint=0, linear=False, return_e2d=False, shards=1):
        """J-Agg / P-Agg.  shards=1: preds [B,K,H,F,17,3]; shards=W: preds [W,B,K,H/W,F,17,3], the rank-major output of
        the all-gather of the per-rank hypothesis shards (distributed.gather_shards)."""
        preds, traj, cam, x2d = self._f32(preds), self._f32(traj), self._f32(cam), self._f32(x2d)
        if preds.dim() == 7:  # rank-major shard buffer (a single shard is the reference layout with a leading 1)
            assert preds.shape[0] == shards
            B, K, H = preds.shape[1], preds.shape[2], preds.shape[3] * shards
        else:
            assert shards == 1
            B, K, H = preds.shape[0], preds.shape[1], preds.shape[2]
        if preds.dim() not in (6, 7) or tuple(preds.shape[-3:]) != (self.frames, 17, 3) or traj.numel() != B * self.frames * 3:
            raise D3dpError(f"jpma: preds {tuple(preds.shape)} / traj {tuple(traj.shape)} do not match F={self.frames}")
        self._expect(x2d, (B, self.frames, 17, 2), "jpma x2d")
        traj = traj.reshape(B, self.frames, 3)
        if cam.dim() == 1:
            cam = cam[None].expand(B, 9).contiguous()
        jagg = torch.empty(B, K, self.frames, 17, 3, dtype=torch.float32, device=self.device)
        pagg = torch.empty_like(jagg)
        idx = torch.empty(B, K, self.frames, 17, dtype=torch.int32, device=self.device)
        e2d = torch.empty(B, K, self.frames, 17, dtype=torch.float32, device=self.device) if return_e2d else None
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_jpma(self.handle, ptr(preds), ptr(traj), ptr(cam), ptr(x2d), ptr(jagg),
                                                  ptr(idx), ptr(pagg), ptr(e2d), B, K, H, int(root_joint), int(linear),
                                                  int(shards), _stream()), "d3dp_jpma")
        return (jagg, idx, pagg, e2d) if return_e2d else (jagg, idx, pagg)

    def jpma_gt(self, preds, traj, cam, x2d, gt, root_joint=0, linear=False):
        """JPMA plus the ground-truth-dependent outputs (per-hypothesis 3-D errors, J-Best pose)."""
        preds, traj, cam, x2d, gt = (self._f32(t) for t in (preds, traj, cam, x2d, gt))
        B, K, H = preds.shape[0], preds.shape[1], preds.shape[2]
        self._expect(preds, (B, K, H, self.frames, 17, 3), "jpma_gt preds")
        self._expect(x2d, (B, self.frames, 17, 2), "jpma_gt x2d")
        self._expect(gt, (B, self.frames, 17, 3), "jpma_gt gt")
        traj = traj.reshape(B, self.frames, 3)
        if cam.dim() == 1:
            cam = cam[None].expand(B, 9).contiguous()
        new = lambda *shape, dt=torch.float32: torch.empty(*shape, dtype=dt, device=self.device)  # noqa: E731
        jagg, pagg, jbest = (new(B, K, self.frames, 17, 3) for _ in range(3))
        idx, e2d, e3d = new(B, K, self.frames, 17, dt=torch.int32), new(B, K, self.frames, 17), new(B, K, H, self.frames, 17)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_jpma_gt(
                self.handle, ptr(preds), ptr(traj), ptr(cam), ptr(x2d), ptr(gt), ptr(jagg), ptr(idx), ptr(pagg),
                ptr(e2d), ptr(e3d), ptr(jbest), B, K, H, int(root_joint), int(linear), 1, _stream()), "d3dp_jpma_gt")
        return {"jagg_pose": jagg, "jagg_idx": idx, "pagg_pose": pagg, "e2d_min": e2d, "e3d": e3d, "jbest_pose": jbest}

    def pmpjpe(self, preds, gt, root_joint=0):
        """Procrustes-aligned per-joint errors (Protocol 2): preds [B,K,H,F,17,3] (or [B,K,F,17,3], e.g. the P-Agg
        pose) vs gt [B,F,17,3] -> same leading shape + [17]."""
        preds, gt = self._f32(preds), self._f32(gt)
        squeeze = preds.dim() == 5
        if squeeze:
            preds = preds[:, :, None]
        B, K, H = preds.shape[:3]
        assert preds.shape[3:] == (self.frames, 17, 3) and gt.shape == (B, self.frames, 17, 3)
        err = torch.empty(B, K, H, self.frames, 17, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_pmpjpe(self.handle, ptr(preds), ptr(gt), ptr(err), B, K, H, int(root_joint),
                                                    _stream()), "d3dp_pmpjpe")
        return err[:, :, 0] if squeeze else err

    def philox_normal(self, B, H, per_bh, seed, h_offset=0, H_total=None, draw=0):
        out = torch.empty(B, H, per_bh, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_philox_normal(self.handle, ptr(out), B, H, per_bh, int(seed), h_offset,
                                                           H if H_total is None else H_total, draw, _stream()),
                  "d3dp_philox_normal")
        return out

    # ------------------------------------------------------------------ kernel-level hooks (tests / profiling)
    def test_gemm(self, mode, a16, w16, bias, x=None, ln_a=None, ln_b=None, tpos=None, F=1):
        M, K = a16.shape
        N = w16.shape[0]
        out16 = torch.empty(M, N if mode < 2 else 512, dtype=torch.float16, device=self.device)
        g_a, b_a, eps_a = ln_a if ln_a is not None else (None, None, 0.0)
        g_b, b_b, eps_b = ln_b if ln_b is not None else (None, None, 0.0)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_test_gemm(
                self.handle, mode, ptr(a16), ptr(w16), ptr(bias), ptr(out16), ptr(x), ptr(g_a), ptr(b_a), eps_a,
                ptr(g_b), ptr(b_b), eps_b, ptr(tpos), F, M, N, K, _stream()), "d3dp_test_gemm")
        return out16

    def test_attn(self, temporal, qkv16, n_streams):
        T = qkv16.shape[0]
        o16 = torch.empty(T, 512, dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            check(self.handle, self.lib.d3dp_test_attn(self.handle, int(temporal), ptr(qkv16), ptr(o16), n_streams,
                                                       _stream()), "d3dp_test_attn")
        return o16
