"""BASELINE config 5: throughput / roofline grid over F in {81,243,351}, H in {1,5,20,80}, K in {1,5,10} (B=4 clips,
flip TTA, Philox noise).  One JSON line per cell: poses/s (B*F/t), hypothesis-poses/s, achieved TFLOP/s per GPU
against the measured bf16 peak.

    python profiles/sweep.py [--quick]                                   # 1 GPU: sampler only (as in round 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/sweep.py [--quick]                                      # 8 GPUs: the cell's H hypotheses PER GPU
        (H_total = 8 H, hypothesis-sharded exactly like bench.py: Philox noise by global hypothesis index, one NCCL
        all-gather of the shards, JPMA on every rank); time = max over ranks, barrier on both sides of every cell.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import f_tok, measured_peaks  # noqa: E402
from d3dp_b200 import D3DP  # noqa: E402
from d3dp_b200.distributed import gather_shards  # noqa: E402
from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, flip_2d, make_args,  # noqa: E402
                                 synthetic_camera, synthetic_pose_estimator_state)


def main():
    quick = "--quick" in sys.argv
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    B = 4
    for F in (81, 243, 351):
        sd = synthetic_pose_estimator_state(F, seed=0)
        g = torch.Generator().manual_seed(1)
        x2d = (0.3 * torch.randn(B, F, 17, 2, generator=g))
        x2d_d, x2d_f = x2d.to(dev), flip_2d(x2d).to(dev)
        traj, cam = (t.to(dev) for t in synthetic_camera(B, F))
        for H in (1, 5, 20, 80):
            for K in (1, 5, 10):
                if quick and (H, K) not in ((1, 1), (1, 10), (5, 5), (20, 10), (80, 10)):
                    continue
                model = D3DP(make_args(F), JL, JR, is_train=False, num_proposals=H, sampling_timesteps=K)
                model.pose_estimator.load_state_dict(sd, strict=True)
                model = model.to(dev).eval()
                eng = model.pose_estimator.engine()

                def call(i):
                    preds = model.ddim_sample_flip(x2d_d, None, input_2d_flip=x2d_f, seed=1 + i, h_offset=rank * H,
                                                   H_total=H * world)
                    if world > 1:
                        return eng.jpma(gather_shards(preds, world), traj, cam, x2d_d, shards=world)
                    return preds

                for i in range(2):
                    call(i)
                reps = 3 if H * K >= 100 else 6
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                s.record()
                for i in range(reps):
                    call(2 + i)
                e.record()
                torch.cuda.synchronize()
                t = torch.tensor([s.elapsed_time(e) / reps * 1e-3], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dist.barrier()
                t = t.item()
                flops = f_tok(F) * B * H * F * 17 * K * 2  # per GPU
                row = {"n_gpus": world, "F": F, "B": B, "H_per_gpu": H, "H_total": H * world, "K": K,
                       "ms": round(t * 1e3, 3), "poses_per_s": round(B * F / t, 1),
                       "hyp_poses_per_s": round(B * H * world * F / t, 1), "tflops_per_gpu": round(flops / t / 1e12, 1),
                       "frac_of_bf16_peak": round(flops / t / 1e12 / peaks["tflops_sustained"], 3)}
                if rank == 0:
                    print(json.dumps(row), flush=True)
                del model, eng
                torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
