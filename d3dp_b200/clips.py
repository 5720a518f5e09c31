"""Host-side clip pipeline around the sampler call (reference: main.py:267-299 `eval_data_prepare`, :646-648 flip
construction, :685-696 batching).  Plain torch indexing on whatever device the sequence lives on — with the sequence
already on the GPU nothing here touches the host, so clips of several sequences can be concatenated to keep the
sampler's batch full."""
import torch


def eval_data_prepare(receptive_field, inputs_2d, inputs_3d=None):
    """Cut one sequence [1, N, 17, C] (or [N, 17, C]) into ceil(N / F) clips of F frames: clip i = frames
    [i*F, (i+1)*F), the last clip = the LAST F frames (overlapping its predecessor); a sequence shorter than F is padded
    by repeating its last frame (main.py:267-299).  Returns [n_clips, F, 17, C] tensors (3-D one None if not given)."""
    F = receptive_field

    def cut(x):
        if x is None:
            return None
        x = x.reshape(-1, *x.shape[-2:])
        n = x.shape[0]
        if n < F:
            x = torch.cat([x, x[-1:].expand(F - n, *x.shape[1:])], dim=0)
            n = F
        n_clips = (n + F - 1) // F
        starts = [i * F for i in range(n_clips - 1)] + [n - F]
        idx = torch.tensor(starts, device=x.device)[:, None] + torch.arange(F, device=x.device)[None, :]
        return x[idx]
    if inputs_3d is not None:
        assert inputs_2d.shape[:-1] == inputs_3d.shape[:-1], "2d and 3d inputs shape must be same!"
    return cut(inputs_2d), cut(inputs_3d)


def image_coordinates(x, w, h):
    """Normalised screen coordinates -> pixels (common/camera.py:14-18), used on the 2-D target of the 3DHP J-Agg
    selection (main_3dhp.py:829)."""
    assert x.shape[-1] == 2
    return (x + torch.tensor([1.0, h / w], dtype=x.dtype, device=x.device)) * (w / 2)


def export_layout_3dhp(clip_poses, n_frames):
    """MATLAB layout of the 3DHP pose export (main_3dhp.py:327-332 pose_post_process, :866-871): per-clip poses
    [n_clips, K, F, 17, 3] of one sequence -> [3, 17, N, K] (the last clip supplies the last F frames)."""
    seq = stitch_clips_last_wins(clip_poses, n_frames)     # [K, N, 17, 3]
    return seq.permute(3, 2, 1, 0).contiguous()


def flip_inputs(inputs_2d, kps_left, kps_right):
    """Test-time-augmentation input (main.py:646-648): negate x, swap left/right key points."""
    out = inputs_2d.clone()
    out[..., 0] *= -1
    out[..., list(kps_left) + list(kps_right), :] = out[..., list(kps_right) + list(kps_left), :]
    return out


def batches(n_clips, batch_size):
    """Slices of at most `batch_size` clips (main.py:685-696)."""
    return [slice(i, min(i + batch_size, n_clips)) for i in range(0, n_clips, batch_size)]


def stitch_clips(clip_out, n_frames):
    """Inverse of eval_data_prepare for per-clip outputs [n_clips, ..., F, 17, C] with the frame axis at -3: returns
    [..., N, 17, C] for the original N frames (the overlap of the last clip is resolved in favour of the earlier clip)."""
    F = clip_out.shape[-3]
    n_clips = clip_out.shape[0]
    if n_frames <= F:
        return clip_out[0][..., :n_frames, :, :]
    body = [clip_out[i] for i in range(n_clips - 1)]
    tail = n_frames - (n_clips - 1) * F
    body.append(clip_out[-1][..., F - tail:, :, :])
    return torch.cat(body, dim=-3)


def stitch_clips_last_wins(clip_out, n_frames):
    """As stitch_clips, but the overlap is taken from the LAST clip, which is what the reference's export does
    (main_3dhp.py:328-330 writes clips in order, then overwrites the last F frames)."""
    F = clip_out.shape[-3]
    n_clips = clip_out.shape[0]
    if n_frames <= F:
        return clip_out[0][..., :n_frames, :, :]
    head = n_frames - F
    full = [clip_out[i] for i in range(n_clips - 1)]
    body = torch.cat(full, dim=-3)[..., :head, :, :]
    return torch.cat([body, clip_out[-1]], dim=-3)
