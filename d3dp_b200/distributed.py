"""Hypothesis-sharded multi-GPU sampling (one process per GPU, torch.distributed over NCCL/NVLink).

Every (clip, hypothesis) pair is an independent chain through all K DDIM steps (the denoiser never mixes batch or
hypothesis: common/mixste.py:230,244), so rank r owns hypotheses [r*H/W, (r+1)*H/W) of every clip, draws its noise
by global hypothesis index (Philox, h_offset) and runs the whole sampler with zero communication.  The only exchange
step of the path is the all-gather of the per-rank [B,K,h,F,17,3] predictions before the J-Agg / P-Agg reduction
(main.py:705-718 needs all hypotheses of a pose together).  The reference's own multi-GPU mechanism is
single-process nn.DataParallel over clips (main.py:242-248); hypothesis sharding replaces it.
"""
import torch
import torch.distributed as dist


def shard_range(H_total, world, rank):
    """Contiguous, balanced hypothesis range of `rank`: (h_offset, h_local). Works when world does not divide H."""
    base, rem = divmod(H_total, world)
    h_local = base + (1 if rank < rem else 0)
    h_offset = rank * base + min(rank, rem)
    return h_offset, h_local


def gather_shards(preds, world, group=None, async_op=False):
    """The path's one collective: all-gather the per-rank [B,K,h,F,17,3] shards (equal h on every rank) into the
    rank-major buffer [world,B,K,h,F,17,3] exactly as `all_gather_into_tensor` lays it down — no re-layout copy.
    `Engine.jpma(..., shards=world)` aggregates straight from this layout (global hypothesis r*h + hl = shard r,
    local hl); `shards_to_reference_layout` materialises the reference's [B,K,H,F,17,3] when a caller wants it.
    async_op=True returns (buffer, work): the collective runs on NCCL's own stream, so the caller can queue the next
    sampler call on the compute stream and `work.wait()` on a side stream before the aggregation."""
    if world == 1:
        return (preds[None], None) if async_op else preds[None]
    preds = preds.contiguous()
    out = torch.empty((world,) + tuple(preds.shape), dtype=preds.dtype, device=preds.device)
    # concatenation form along dim 0 (accepted by every backend); `flat` is a view of `out`
    flat = out.view((world * preds.shape[0],) + tuple(preds.shape[1:]))
    work = dist.all_gather_into_tensor(flat, preds, group=group, async_op=async_op)
    return (out, work) if async_op else out


def shards_to_reference_layout(shards):
    """[world,B,K,h,F,17,3] -> the reference's [B,K,world*h,F,17,3] (one permute copy)."""
    W, B, K, h = shards.shape[:4]
    return shards.permute(1, 2, 0, 3, 4, 5, 6).reshape(B, K, W * h, *shards.shape[4:])


def gather_hypotheses(preds, world, group=None):
    """all-gather [B,K,h,F,17,3] from every rank into [B,K,h*world,F,17,3] (equal h on all ranks), ordered by rank =
    ordered by global hypothesis index.  One collective: dist.all_gather_into_tensor."""
    if world == 1:
        return preds
    return shards_to_reference_layout(gather_shards(preds, world, group))


def gather_hypotheses_uneven(preds, H_total, world, rank, group=None):
    """Same for H_total not divisible by world: pad every shard to the largest, gather, drop the padding."""
    if world == 1:
        return preds
    h_max = -(-H_total // world)
    B, K, h = preds.shape[:3]
    if h < h_max:
        pad = torch.zeros((B, K, h_max - h) + tuple(preds.shape[3:]), dtype=preds.dtype, device=preds.device)
        preds = torch.cat([preds, pad], dim=2)
    full = gather_hypotheses(preds, world, group).reshape(B, K, world, h_max, *preds.shape[3:])
    parts = [full[:, :, r, :shard_range(H_total, world, r)[1]] for r in range(world)]
    return torch.cat(parts, dim=2)


def sample_sharded(sampler, x2d, x2d_flip, H_total, seed, rank=None, world=None, group=None):
    """Run `sampler(x2d, x2d_flip, h_local, h_offset, H_total, seed) -> [B,K,h_local,F,17,3]` on this rank's
    hypothesis shard and return the gathered [B,K,H_total,F,17,3] (identical on every rank).
    `sampler` is normally a closure over a D3DP module (see bench.py); the host logic is backend-agnostic, which is
    what the gloo world_size=2 CPU tests exercise with the oracle as the sampler."""
    world = dist.get_world_size(group) if world is None else world
    rank = dist.get_rank(group) if rank is None else rank
    h_offset, h_local = shard_range(H_total, world, rank)
    preds = sampler(x2d, x2d_flip, h_local, h_offset, H_total, seed)
    if H_total % world == 0:
        return gather_hypotheses(preds, world, group)
    return gather_hypotheses_uneven(preds, H_total, world, rank, group)
