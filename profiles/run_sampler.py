"""One sampler call at the bench width (F=243, B=4, H=20, flip) with K steps (default 1), for ncu captures of the
kernels that only exist inside the sampler (embed, head, time-MLP, DDIM step):
    D3DP_GRAPH=0 ncu --metrics gpu__time_duration.sum -k regex:embed_kernel python profiles/run_sampler.py [K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3dp_b200 import D3DP  # noqa: E402
from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, flip_2d, make_args,  # noqa: E402
                                 synthetic_pose_estimator_state)

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1
F, B, H = 243, 4, 20
model = D3DP(make_args(F), JL, JR, is_train=False, num_proposals=H, sampling_timesteps=K)
model.pose_estimator.load_state_dict(synthetic_pose_estimator_state(F, seed=0), strict=True)
model = model.cuda().eval()
x2d = 0.3 * torch.randn(B, F, 17, 2, generator=torch.Generator().manual_seed(1234))
out = model.ddim_sample_flip(x2d.cuda(), None, input_2d_flip=flip_2d(x2d).cuda(), seed=7)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().max()))
