"""Deterministic synthetic weights and inputs (no checkpoint or dataset is available offline).

Every tensor is generated from a torch CPU generator seeded by crc32(key) ^ seed, so the reference (when golden
vectors are produced), the oracle, the tests and bench.py all see bit-identical values on any machine without
shipping 140 MB of weights.  Shapes/keys are the reference's `pose_estimator.*` state_dict
(common/mixste.py:142-210; 208 tensors).
"""
import math
import types
import zlib

import torch

H36M_JOINTS_LEFT = [4, 5, 6, 11, 12, 13]
H36M_JOINTS_RIGHT = [1, 2, 3, 14, 15, 16]
# Human3.6M camera 0 intrinsics, normalised (common/h36m_dataset.py:20-29,216-231): f(2) c(2) k(3) p(2)
H36M_CAM0 = [2.2901, 2.2876, 0.0251, 0.0289, -0.2071, 0.2478, -0.0031, -0.00098, -0.00142]


def make_args(frames, scale=1.0, depth=8, flip=True):
    """The fields D3DP.__init__ reads from the reference's argparse namespace (common/arguments.py:49-50,58,101-102,112)."""
    return types.SimpleNamespace(number_of_frames=frames, test_time_augmentation=flip, timestep=1000, scale=scale,
                                 cs=512, dep=depth)


def pose_estimator_shapes(frames, depth=8, C=512):
    s = {
        "Spatial_patch_to_embedding.weight": (C, 5), "Spatial_patch_to_embedding.bias": (C,),
        "Spatial_pos_embed": (1, 17, C), "Temporal_pos_embed": (1, frames, C),
        "time_mlp.1.weight": (2 * C, C), "time_mlp.1.bias": (2 * C,),
        "time_mlp.3.weight": (C, 2 * C), "time_mlp.3.bias": (C,),
    }
    for kind in ("STEblocks", "TTEblocks"):
        for d in range(depth):
            p = f"{kind}.{d}."
            s.update({
                p + "norm1.weight": (C,), p + "norm1.bias": (C,),
                p + "attn.qkv.weight": (3 * C, C), p + "attn.qkv.bias": (3 * C,),
                p + "attn.proj.weight": (C, C), p + "attn.proj.bias": (C,),
                p + "norm2.weight": (C,), p + "norm2.bias": (C,),
                p + "mlp.fc1.weight": (2 * C, C), p + "mlp.fc1.bias": (2 * C,),
                p + "mlp.fc2.weight": (C, 2 * C), p + "mlp.fc2.bias": (C,),
            })
    s.update({
        "Spatial_norm.weight": (C,), "Spatial_norm.bias": (C,),
        "Temporal_norm.weight": (C,), "Temporal_norm.bias": (C,),
        "head.0.weight": (C,), "head.0.bias": (C,), "head.1.weight": (3, C), "head.1.bias": (3,),
    })
    return s


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synthetic_pose_estimator_state(frames, depth=8, seed=0):
    """{key: float32 tensor}: Linear ~ U(+-1/sqrt(fan_in)) like nn.Linear's default init, LayerNorm near (1, 0) but
    not exactly (so gamma/beta are exercised), positional embeddings N(0, 0.02) (the reference zero-inits them)."""
    out = {}
    for key, shape in pose_estimator_shapes(frames, depth).items():
        g = _gen(key, seed)
        if "pos_embed" in key:
            t = 0.02 * torch.randn(shape, generator=g)
        elif "norm" in key or key.startswith("head.0"):
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if key.endswith("weight") \
                else 0.05 * torch.randn(shape, generator=g)
        else:
            if key.endswith("weight"):
                fan_in = shape[1]
            else:
                fan_in = {"Spatial_patch_to_embedding.bias": 5, "time_mlp.1.bias": 512, "time_mlp.3.bias": 1024,
                          "head.1.bias": 512}.get(key)
                if fan_in is None:
                    fan_in = 1024 if key.endswith("mlp.fc2.bias") else 512
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[key] = t.float()
    return out


def flip_2d(x2d, joints_left=H36M_JOINTS_LEFT, joints_right=H36M_JOINTS_RIGHT):
    """Test-time-augmentation input (main.py:646-648): negate x, swap left/right keypoints."""
    f = x2d.clone()
    f[..., 0] *= -1
    f[..., joints_left + joints_right, :] = f[..., joints_right + joints_left, :]
    return f


def synthetic_inputs(B, H, K, frames, seed=1234, noise_seed=123):
    """x2d ~ 0.3 N(0,1) in normalised screen coordinates, its flip, and the injected sampler noise."""
    g = torch.Generator().manual_seed(seed)
    x2d = 0.3 * torch.randn(B, frames, 17, 2, generator=g)
    gn = torch.Generator().manual_seed(noise_seed)
    noise_init = torch.randn(B, H, frames, 17, 3, generator=gn)
    noise_steps = torch.randn(max(K - 1, 0), B, H, frames, 17, 3, generator=gn)
    return x2d, flip_2d(x2d), noise_init, noise_steps


def synthetic_camera(B, frames, seed=99):
    """Root trajectory with z > 0 (so X/Z is well inside the clamp) and the H36M camera-0 intrinsics."""
    g = torch.Generator().manual_seed(seed)
    traj = torch.tensor([0.0, 0.0, 5.0]) + 0.5 * torch.randn(B, frames, 1, 3, generator=g)
    cam = torch.tensor(H36M_CAM0, dtype=torch.float32)[None].repeat(B, 1)
    return traj, cam
