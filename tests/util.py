"""Shared helpers for the tests (test infrastructure)."""
import os

import torch

from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, make_args,  # noqa: F401
                                 synthetic_inputs, synthetic_pose_estimator_state)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["f27_flip", "f27_noflip_k1", "f27_scale2", "f9_depth2", "f243_flip",
                "f243_k10",   # K = 10, the benchmarked depth of the DDIM loop (one chain)
                "f243_c2"]    # BASELINE config 2 exactly: F=243, B=1, H=5, K=5


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=True)


def case_inputs(case):
    sd = synthetic_pose_estimator_state(case["F"], depth=case["depth"], seed=case["weight_seed"])
    x2d, x2d_flip, n0, ns = synthetic_inputs(case["B"], case["H"], case["K"], case["F"])
    return sd, x2d, x2d_flip, n0, ns


def build_model(F, H, K, sd, scale=1.0, depth=8, flip=True, device="cuda"):
    from d3dp_b200 import D3DP
    model = D3DP(make_args(F, scale, depth, flip), JL, JR, is_train=False, num_proposals=H, sampling_timesteps=K)
    model.pose_estimator.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def mpjpe_distance(a, b):
    d = torch.norm(a.double().cpu() - b.double().cpu(), dim=-1)
    return d.mean().item(), d.max().item()
