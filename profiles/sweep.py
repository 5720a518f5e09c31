"""BASELINE config 5: throughput / roofline grid over F in {81,243,351}, H in {1,5,20,80}, K in {1,5,10} (B=4 clips,
flip TTA, Philox noise, JPMA excluded).  Prints one JSON line per cell: poses/s (B*F/t), hypothesis-poses/s,
achieved TFLOP/s against the measured bf16 peak.   usage: python profiles/sweep.py [--quick]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import f_tok, measured_peaks  # noqa: E402
from d3dp_b200 import D3DP  # noqa: E402
from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, flip_2d,  # noqa: E402
                                 synthetic_pose_estimator_state)
from d3dp_b200.synthetic import make_args  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    peaks = measured_peaks()
    B = 4
    rows = []
    for F in (81, 243, 351):
        sd = synthetic_pose_estimator_state(F, seed=0)
        g = torch.Generator().manual_seed(1)
        x2d = (0.3 * torch.randn(B, F, 17, 2, generator=g))
        x2d_d, x2d_f = x2d.cuda(), flip_2d(x2d).cuda()
        for H in (1, 5, 20, 80):
            for K in (1, 5, 10):
                if quick and (H, K) not in ((1, 1), (20, 10)):
                    continue
                model = D3DP(make_args(F), JL, JR, is_train=False, num_proposals=H, sampling_timesteps=K)
                model.pose_estimator.load_state_dict(sd, strict=True)
                model = model.cuda().eval()
                for _ in range(2):
                    model.ddim_sample_flip(x2d_d, None, input_2d_flip=x2d_f, seed=1)
                reps = 3 if H * K >= 100 else 6
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                s.record()
                for i in range(reps):
                    model.ddim_sample_flip(x2d_d, None, input_2d_flip=x2d_f, seed=2 + i)
                e.record()
                torch.cuda.synchronize()
                t = s.elapsed_time(e) / reps * 1e-3
                flops = f_tok(F) * B * H * F * 17 * K * 2
                row = {"F": F, "B": B, "H": H, "K": K, "ms": round(t * 1e3, 3), "poses_per_s": round(B * F / t, 1),
                       "hyp_poses_per_s": round(B * H * F / t, 1), "tflops": round(flops / t / 1e12, 1),
                       "frac_of_bf16_peak": round(flops / t / 1e12 / peaks["tflops_sustained"], 3)}
                rows.append(row)
                print(json.dumps(row), flush=True)
                del model
                torch.cuda.empty_cache()
    return rows


if __name__ == "__main__":
    main()
