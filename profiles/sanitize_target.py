"""Target process of the compute-sanitizer runs (profiles/r02_sanitizer.sh): every hot kernel once at the smoke shape
(F=27) and at the bench's sequence length (F=243, one clip / one hypothesis), through the public API.
    compute-sanitizer --tool racecheck --kernel-regex kns=d3dp python profiles/sanitize_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3dp_b200 import D3DP  # noqa: E402
from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, make_args, synthetic_camera,  # noqa: E402
                                 synthetic_inputs, synthetic_pose_estimator_state)

shapes = [(27, 1, 2, 2)] + ([(243, 1, 1, 1)] if os.environ.get("SAN_FULL", "1") == "1" else [])
for F, B, H, K in shapes:
    depth = int(os.environ.get("SAN_DEPTH", "2"))
    sd = synthetic_pose_estimator_state(F, depth=depth, seed=0)
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H, K, F)
    model = D3DP(make_args(F, depth=depth), JL, JR, is_train=False, num_proposals=H, sampling_timesteps=K)
    model.pose_estimator.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    out = model.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=x2d_flip.cuda(), seed=3)
    lst = model.ddim_sample(x2d.cuda(), None, noise_init=n0, noise_steps=ns)
    eng = model.pose_estimator.engine()
    traj, cam = synthetic_camera(B, F)
    jagg, idx, pagg = eng.jpma(out, traj, cam, x2d)
    gt = 0.4 * torch.randn(B, F, 17, 3)
    eng.jpma_gt(out, traj, cam, x2d, gt)
    eng.pmpjpe(out, gt)
    eng.q_sample(gt, torch.randn_like(gt), torch.randint(0, 1000, (B,)), clamp=True)
    masks = model.pose_estimator.draw_drop_masks(B * H, "cuda")
    model.pose_estimator(x2d.cuda(), n0.cuda(), torch.full((B,), 500), drop_masks=masks)
    torch.cuda.synchronize()
    print(f"F={F} B={B} H={H} K={K} depth={depth}: ok, |out|max {out.abs().max().item():.3f}", flush=True)
