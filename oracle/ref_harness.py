"""Run the UNMODIFIED reference (/root/reference, or $D3DP_REF) on the CPU with injected noise — TEST INFRASTRUCTURE.

Only usable where the reference tree exists (this build container); it is what pins `oracle/d3dp_oracle.py` and what
generated `tests/golden/*.pt`.  Shims: a `timm` stub (only DropPath is used, Identity at eval),
a proxy for the module-global `torch` inside common.diffusionpose whose randn/randn_like serve the injected noise and
drop `device='cuda'`, and `Tensor.cuda` as identity.  No reference source is copied: it is imported where it lies.
"""
import os
import sys
import types

import torch

REF = os.environ.get("D3DP_REF", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF, "common", "diffusionpose.py"))


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m


class _DropPath(torch.nn.Module):
    """Stand-in for timm.models.layers.DropPath (timm is not installed; its drop_path is: identity unless training
    and p > 0, else x * bernoulli(1-p)/(1-p) drawn per sample of the FIRST axis, shape (N,1,...,1)).  The factor
    tensors are served from `queue` (filled by the caller, in execution order) so that the unmodified reference's
    training-mode forward can be compared with the oracle on identical draws."""
    queue = []

    def __init__(self, p=0.0):
        super().__init__()
        self.drop_prob = p

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        f = _DropPath.queue.pop(0)
        assert f.numel() == x.shape[0], (f.shape, x.shape)
        return x * f.reshape((x.shape[0],) + (1,) * (x.dim() - 1)).to(x.dtype)


class _TorchProxy:
    """Stands in for `torch` inside common/diffusionpose.py: serves queued noise, ignores device='cuda'."""

    def __init__(self, queue):
        self._q = list(queue)

    def randn(self, shape, device=None, **kw):
        t = self._q.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t.clone()

    def randn_like(self, x):
        t = self._q.pop(0)
        assert t.shape == x.shape
        return t.clone()

    def full(self, size, fill, device=None, dtype=None):
        return torch.full(size, fill, dtype=dtype)

    def __getattr__(self, name):
        return getattr(torch, name)


def import_reference():
    if "timm" not in sys.modules:
        _stub("timm")
        _stub("timm.data", IMAGENET_DEFAULT_MEAN=None, IMAGENET_DEFAULT_STD=None)
        _stub("timm.models")
        _stub("timm.models.helpers", load_pretrained=None)
        _stub("timm.models.layers", DropPath=_DropPath, to_2tuple=None, trunc_normal_=None)
        _stub("timm.models.registry", register_model=None)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import common.diffusionpose as dp
    return dp


def make_args(frames, scale=1.0, depth=8):
    return types.SimpleNamespace(number_of_frames=frames, test_time_augmentation=True, timestep=1000, scale=scale,
                                 cs=512, dep=depth)


def build_reference_model(frames, H, K, state, joints_left, joints_right, scale=1.0, depth=8, flip=True):
    dp = import_reference()
    args = make_args(frames, scale, depth)
    args.test_time_augmentation = flip
    model = dp.D3DP(args, joints_left, joints_right, is_train=False, num_proposals=H, sampling_timesteps=K).eval()
    missing = model.pose_estimator.load_state_dict(state, strict=True)
    model.device = torch.device("cpu")  # ddim_sample reads self.device (reference bug)
    return model


def run_reference_train_forward(frames, state, x2d, x_t, t, drop_masks, depth=8):
    """The reference's MixSTE2 in TRAINING mode (is_train=True layout, DropPath 0.1 active) with injected DropPath
    factors: `drop_masks` is the depth x 4 list of d3dp_b200.MixSTE2.draw_drop_masks; blocks whose rate is 0 are
    nn.Identity in the reference (common/mixste.py:100) and consume nothing."""
    dp = import_reference()
    model = dp.D3DP(make_args(frames, depth=depth), [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16], is_train=True)
    model.pose_estimator.load_state_dict(state, strict=True)
    model.train()
    rates = [x.item() for x in torch.linspace(0, 0.1, depth)]
    _DropPath.queue = [m.reshape(-1) for d in range(depth) for m in drop_masks[4 * d:4 * d + 4] if rates[d] > 0]
    with torch.no_grad():
        out = model.pose_estimator(x2d, x_t, t)
    assert not _DropPath.queue
    return out


def run_reference_sampler(model, x2d, x2d_flip, noise_init, noise_steps):
    """model.forward with the noise draws injected; returns [B,K,H,F,17,3] (flip path) on the CPU."""
    dp = import_reference()
    queue = [noise_init] + [noise_steps[i] for i in range(noise_steps.shape[0])]
    real_torch, real_cuda = dp.torch, torch.Tensor.cuda
    dp.torch = _TorchProxy(queue)
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            out = model(x2d, None, input_2d_flip=x2d_flip)
    finally:
        dp.torch = real_torch
        torch.Tensor.cuda = real_cuda
    if isinstance(out, list):
        out = torch.stack(out, dim=1)
    return out
