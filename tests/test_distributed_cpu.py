"""world_size=2 gloo test (CPU) of the hypothesis-sharded sampling path: each rank runs the sampler on its
hypothesis shard and the shards are all-gathered along the hypothesis axis (d3dp_b200/distributed.py).  The per-rank
sampler here is the CPU oracle (the CUDA sampler cannot run without a GPU); the host logic under test — shard
arithmetic, gather layout, ordering by global hypothesis index — is the same code bench.py runs over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, synthetic_inputs,
                                 synthetic_pose_estimator_state)

F, B, K, DEPTH = 9, 2, 2, 1


def _sampler_factory(sd, n0, ns):
    from oracle import d3dp_oracle as orc

    def sampler(x2d, x2d_flip, h_local, h_offset, H_total, seed):
        sl = slice(h_offset, h_offset + h_local)
        with torch.no_grad():
            return orc.ddim_sample(sd, x2d, x2d_flip, h_local, K, n0[:, sl], ns[:, :, sl], JL, JR, depth=DEPTH)
    return sampler


def _worker(rank, world, port, H_total, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from d3dp_b200.distributed import sample_sharded
    sd = synthetic_pose_estimator_state(F, depth=DEPTH, seed=5)
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H_total, K, F)
    full = sample_sharded(_sampler_factory(sd, n0, ns), x2d, x2d_flip, H_total, seed=0)
    if rank == 0:
        torch.save(full, out_path)
    # every rank holds the same gathered tensor
    ref = [torch.empty_like(full) for _ in range(world)]
    dist.all_gather(ref, full)
    assert all(torch.equal(r, full) for r in ref)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("H_total", [4, 5])
def test_sharded_sampling_matches_single_process(tmp_path, H_total):
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, _free_port(), H_total, out), nprocs=2, join=True)
    gathered = torch.load(out, weights_only=True)
    sd = synthetic_pose_estimator_state(F, depth=DEPTH, seed=5)
    x2d, x2d_flip, n0, ns = synthetic_inputs(B, H_total, K, F)
    single = _sampler_factory(sd, n0, ns)(x2d, x2d_flip, H_total, 0, H_total, 0)
    assert gathered.shape == (B, K, H_total, F, 17, 3)
    assert torch.equal(gathered, single)  # hypotheses are independent chains: sharding is exact
