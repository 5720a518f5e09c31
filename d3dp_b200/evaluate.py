"""Sequence-level evaluation around the sampler (reference: main.py:634-760 `evaluate`, main_3dhp.py:735-912): cut
every test sequence into clips, build the flip input, run the sampler, aggregate hypotheses (JPMA) and accumulate the
per-step errors weighted by frames.

What differs from the reference's loop, and why:
  * every sequence is uploaded once and cut / flipped ON THE DEVICE (`clips.py`), so there is no per-sequence
    numpy -> torch -> cuda hop inside the loop;
  * with `packed=True` clips of different sequences share sampler batches, so B is always `batch_size` (the reference
    runs a short last batch per sequence, main.py:688-696).  The logged numbers do not change: J-Best / P-Agg / J-Agg
    are means over (clip, frame, joint), and P-Best — a min over hypotheses of a per-batch mean — is re-assembled here
    per REFERENCE batch (sequence, batch_cnt) from per-clip sums, then weighted by that batch's frames exactly as
    main.py:720-724 does;
  * errors come from the fused kernels (`metrics.py`); nothing is synchronised with the host until the final sums;
  * the clip preparation runs on its own CUDA stream, one batch ahead of the sampler (`overlap=True`): the upload of
    the (pinned) sequences, the cutting / flipping of all clips and the gather of batch i+1's inputs are queued before
    batch i's sampler graph is launched on the compute stream and run underneath it; the compute stream waits on
    a per-batch event only.  The reference does all of this on the host between two sampler calls (main.py:646-696).
"""
import contextlib

import torch

from . import clips as C
from .metrics import jpma_metrics

P1_KEYS = ("J-Best", "P-Best", "P-Agg", "J-Agg")


def evaluate_sequences(model, sequences, kps_left, kps_right, batch_size, root_joint=0, linear=False, protocol2=False,
                       packed=True, seed=None, sampler=None, return_poses=False, overlap=True):
    """`sequences`: list of dicts {"x2d": [N,17,2], "gt": [N,17,3] camera-space poses, "cam": [9] intrinsics}.
    `model`: a d3dp_b200.D3DP on the GPU in eval mode.  Returns {"J-Best","P-Best","P-Agg","J-Agg": [K] tensors}
    (+ "P2-*" with protocol2=True, + per-sequence stitched "jagg_pose"/"pagg_pose" lists with return_poses=True).
    `sampler(x2d, x2d_flip, batch_index)` may replace the model call (tests)."""
    eng = model.pose_estimator.engine()
    dev, F = eng.device, model.frames
    cuda = dev.type == "cuda"  # (the host-logic tests drive this function with a CPU stand-in engine: no streams there)
    main = torch.cuda.current_stream(dev) if cuda else None
    prep = torch.cuda.Stream(device=dev) if (cuda and overlap) else main
    on_prep = (lambda: torch.cuda.stream(prep)) if cuda else contextlib.nullcontext
    if cuda:
        prep.wait_stream(main)  # inputs that already live on the device were produced on the compute stream

    def upload(t):  # host tensors go up asynchronously from pinned memory, on the preparation stream
        if cuda and t.device.type == "cpu":
            t = t.to(torch.float32).pin_memory()
        return t.to(dev, torch.float32, non_blocking=True)
    with on_prep():
        prepared = _cut_all(sequences, upload, F, kps_left, kps_right, batch_size, root_joint, dev)
    x2d_c, flip_c, gt_c, traj_c, cam_c, group, nframes, n_groups = prepared
    n_clips = x2d_c.shape[0]

    with on_prep():
        if packed:
            batch_list = [torch.arange(s.start, s.stop, device=dev) for s in C.batches(n_clips, batch_size)]
        else:
            batch_list = [torch.nonzero(group == g_).flatten() for g_ in range(n_groups)]

    def prepare(bi):
        """Gather the inputs of batch bi on the preparation stream; returns them with the event the compute stream
        waits on.  Called one batch ahead, i.e. before the previous batch's sampler has been launched."""
        idx = batch_list[bi]
        with on_prep():
            items = [x2d_c[idx].contiguous(), flip_c[idx].contiguous(), gt_c[idx], traj_c[idx], cam_c[idx].contiguous(),
                     group[idx]]
            ev = None
            if cuda:
                ev = torch.cuda.Event()
                ev.record(prep)
        if cuda:
            for t in items:
                t.record_stream(main)
        return items, ev

    K = H = None
    acc, poses = {}, {"jagg_pose": [], "pagg_pose": []}
    nxt = prepare(0)
    for bi in range(len(batch_list)):
        (xb, fb, gtb, trajb, camb, grp), ev = nxt
        if bi + 1 < len(batch_list):
            nxt = prepare(bi + 1)
        if cuda:
            main.wait_event(ev)
        if sampler is not None:
            preds = sampler(xb, fb, bi)
        else:
            preds = model.ddim_sample_flip(xb, None, input_2d_flip=fb, seed=None if seed is None else seed + bi)
        m = jpma_metrics(eng, preds, gtb, trajb, camb, xb, root_joint=root_joint, linear=linear, protocol2=protocol2)
        if K is None:
            K, H = preds.shape[1], preds.shape[2]
            shapes = {"jbest": (K,), "per_h": (K, H), "pagg": (K,), "jagg": (K,)}
            if protocol2:
                shapes.update({"p2_" + k: v for k, v in shapes.items()})
            acc = {k: torch.zeros((n_groups,) + v, device=dev, dtype=torch.float64) for k, v in shapes.items()}
        sel = m["jagg_idx"].long().unsqueeze(2)
        sums = {"jbest": m["e3d"].min(dim=2).values, "per_h": m["e3d"],
                "pagg": torch.norm(m["pagg_pose"] - gtb[:, None], dim=-1),
                "jagg": torch.gather(m["e3d"], 2, sel).squeeze(2)}
        if protocol2:
            sums.update({"p2_jbest": m["pe3d"].min(dim=2).values, "p2_per_h": m["pe3d"], "p2_pagg": m["pe3d_mean"],
                         "p2_jagg": torch.gather(m["pe3d"], 2, sel).squeeze(2)})
        for k, v in sums.items():
            acc[k].index_add_(0, grp, v.sum((-1, -2)).double())            # per-clip sums -> their reference batch
        if return_poses:
            poses["jagg_pose"].append(m["jagg_pose"])
            poses["pagg_pose"].append(m["pagg_pose"])
    if cuda:
        main.wait_stream(prep)
    return _finish(acc, poses, group, batch_list, nframes, n_groups, n_clips, F, K, H, protocol2, return_poses, dev)


def _cut_all(sequences, upload, F, kps_left, kps_right, batch_size, root_joint, dev):
    """Upload every sequence once and cut / flip all of its clips on the device (clips.py)."""
    x2d_c, flip_c, gt_c, traj_c, cam_c, group, nframes = [], [], [], [], [], [], []
    n_groups = 0
    for si, s in enumerate(sequences):
        x2d = upload(s["x2d"]).reshape(-1, 17, 2)
        gt = upload(s["gt"]).reshape(-1, 17, 3)
        nframes.append(x2d.shape[0])
        a, g = C.eval_data_prepare(F, x2d, gt)
        b, _ = C.eval_data_prepare(F, C.flip_inputs(x2d, kps_left, kps_right))
        traj_c.append(g[:, :, root_joint:root_joint + 1].clone())          # main.py:682 / main_3dhp.py:771
        g = g.clone()
        g[:, :, root_joint] = 0                                            # main.py:683
        n = a.shape[0]
        x2d_c.append(a); flip_c.append(b); gt_c.append(g)
        cam_c.append(upload(s["cam"]).reshape(1, 9).expand(n, 9))
        group.append(n_groups + torch.arange(n, device=dev) // batch_size)  # the reference's (sequence, batch_cnt)
        n_groups += (n + batch_size - 1) // batch_size
    x2d_c, flip_c, gt_c, traj_c = (torch.cat(t) for t in (x2d_c, flip_c, gt_c, traj_c))
    cam_c, group = torch.cat(cam_c).contiguous(), torch.cat(group)
    return x2d_c, flip_c, gt_c, traj_c, cam_c, group, nframes, n_groups



def _finish(acc, poses, group, batch_list, nframes, n_groups, n_clips, F, K, H, protocol2, return_poses, dev):
    """Frame-weighted means of the per-reference-batch errors (main.py:720-724) and the stitched poses."""
    clips_per_group = torch.zeros(n_groups, device=dev, dtype=torch.float64).index_add_(
        0, group, torch.ones(n_clips, device=dev, dtype=torch.float64))
    frames_w = clips_per_group * F                                         # main.py:720 weight = B * F of that batch
    elems = (frames_w * 17)[:, None]
    total = frames_w.sum()

    def weighted(per_group_mean):                                          # sum_g w_g * e_g / sum_g w_g
        return ((per_group_mean * frames_w[:, None]).sum(0) / total).float()

    out = {}
    for prefix, names in (("", P1_KEYS),) + ((("p2_", tuple("P2-" + k for k in P1_KEYS)),) if protocol2 else ()):
        out[names[0]] = weighted(acc[prefix + "jbest"] / elems)
        out[names[1]] = weighted((acc[prefix + "per_h"] / elems[:, :, None]).min(dim=2).values)
        out[names[2]] = weighted(acc[prefix + "pagg"] / elems)
        out[names[3]] = weighted(acc[prefix + "jagg"] / elems)
    if return_poses:
        order = torch.cat(batch_list)
        inv = torch.empty_like(order)
        inv[order] = torch.arange(n_clips, device=dev)
        for k in poses:
            allp = torch.cat(poses[k])[inv]                                # back to clip order
            per_seq, o = [], 0
            for n in nframes:
                nc = max((n + F - 1) // F, 1)
                per_seq.append(C.stitch_clips_last_wins(allp[o:o + nc], n))  # [K, N, 17, 3]
                o += nc
            out[k] = per_seq
    out["n_clips"], out["n_batches"] = n_clips, len(batch_list)
    return out
