"""d3dp_b200 — B200-native (sm_100a) implementation of the D3DP diffusion-sampling hot path.

Public surface mirrors the reference: `D3DP(args, joints_left, joints_right, is_train, num_proposals,
sampling_timesteps)` (common/diffusionpose.py:55) with `.forward / .ddim_sample / .ddim_sample_flip`, plus the JPMA
aggregation (`jpma`) and the hypothesis-sharded multi-GPU sampler (`distributed`).
"""
__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not need torch.cuda
    if name == "D3DP":
        from .diffusionpose import D3DP
        return D3DP
    if name == "MixSTE2":
        from .mixste import MixSTE2
        return MixSTE2
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)
