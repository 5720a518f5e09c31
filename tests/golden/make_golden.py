"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (/root/reference) on the CPU:
    python tests/golden/make_golden.py
Weights/inputs/noise are regenerated from seeds (d3dp_b200/synthetic.py), so only the reference OUTPUTS are stored.
Each case records the generating arguments; tests rebuild the inputs from them."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, synthetic_inputs,  # noqa: E402
                                 synthetic_pose_estimator_state)
from oracle import ref_harness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name,           F,  B, H, K, flip, scale, depth, weight_seed
    ("f27_flip",      27, 2, 3, 4, True, 1.0, 8, 0),
    ("f27_noflip_k1", 27, 2, 1, 1, False, 1.0, 8, 0),   # BASELINE config 1 (D3DP.ddim_sample smoke)
    ("f27_scale2",    27, 1, 2, 3, True, 2.0, 8, 1),
    ("f9_depth2",     9,  2, 2, 2, True, 1.0, 2, 2),
    ("f243_flip",     243, 1, 2, 2, True, 1.0, 8, 0),
    # round 2: the benchmarked depth of the DDIM loop, and BASELINE config 2 exactly as stated
    ("f243_k10",      243, 1, 1, 10, True, 1.0, 8, 0),  # K=10 (the headline setting's step count), one chain
    ("f243_c2",       243, 1, 5, 5, True, 1.0, 8, 0),   # BASELINE config 2: F=243, B=1, H=5, K=5
]


def main():
    assert rh.available()
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    for name, F, B, H, K, flip, scale, depth, wseed in CASES:
        if only and name not in only:
            continue
        sd = synthetic_pose_estimator_state(F, depth=depth, seed=wseed)
        x2d, x2d_flip, n0, ns = synthetic_inputs(B, H, K, F)
        model = rh.build_reference_model(F, H, K, sd, JL, JR, scale=scale, depth=depth, flip=flip)
        out = rh.run_reference_sampler(model, x2d, x2d_flip if flip else None, n0, ns)
        case = dict(name=name, F=F, B=B, H=H, K=K, flip=flip, scale=scale, depth=depth, weight_seed=wseed,
                    preds=out.float().clone())
        if name == "f27_flip":  # also pin one bare denoiser forward (MixSTE2.forward)
            t = torch.tensor([499, 37], dtype=torch.long)
            with torch.no_grad():
                case["denoise_t"] = t
                case["denoise_out"] = model.pose_estimator(x2d, n0.clamp(-1.1, 1.1), t).float().clone()
        torch.save(case, os.path.join(HERE, name + ".pt"))
        print(name, tuple(out.shape), f"|x|max {out.abs().max():.3f}")


if __name__ == "__main__":
    main()
