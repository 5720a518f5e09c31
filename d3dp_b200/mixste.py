"""`MixSTE2` with the reference's constructor, parameter names and forward signature (common/mixste.py:141-298), whose
forward pass runs in the sm_100a kernels behind the C ABI instead of ATen.

The module tree below only *holds* the parameters (so `.state_dict()`, `.load_state_dict(strict=True)`, `.cuda()`,
`.parameters()` and `nn.DataParallel` behave exactly as with the reference); it is built in the reference's
registration order so `torch.manual_seed(s)` + construction yields the same default initialisation.
"""
from functools import partial

import torch
import torch.nn as nn

from .engine import Engine, H36M_JOINTS_LEFT, H36M_JOINTS_RIGHT


class _Attention(nn.Module):  # parameter holder for common/mixste.py:46-61
    def __init__(self, dim, qkv_bias=True):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):  # parameter holder for common/mixste.py:24-35
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):  # parameter holder for common/mixste.py:84-111
    def __init__(self, dim, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class MixSTE2(nn.Module):
    def __init__(self, num_frame=9, num_joints=17, in_chans=2, embed_dim_ratio=32, depth=4, num_heads=8, mlp_ratio=2.,
                 qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=None,
                 is_train=True, joints_left=H36M_JOINTS_LEFT, joints_right=H36M_JOINTS_RIGHT, scale=1.0,
                 output_scale=1.0):
        super().__init__()
        if (num_joints, in_chans, embed_dim_ratio, num_heads, float(mlp_ratio), bool(qkv_bias), qk_scale) != \
                (17, 2, 512, 8, 2.0, True, None):
            raise ValueError("d3dp_b200 implements the D3DP configuration of MixSTE2 only: 17 joints, 2 input "
                             "channels, embed_dim_ratio=512, 8 heads, mlp_ratio=2, qkv_bias=True, qk_scale=None")
        if not 1 <= depth <= 8 or not 1 <= num_frame <= 384:
            raise ValueError("d3dp_b200 supports depth 1..8 and 1..384 frames in this build")
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        C = embed_dim_ratio
        self.is_train = is_train
        self.num_frame, self.block_depth = num_frame, depth
        # stochastic depth decay rule (common/mixste.py:186): block i of both stacks drops a branch with dpr[i]
        self.drop_path_rates = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self._joints_left, self._joints_right, self._scale = list(joints_left), list(joints_right), float(scale)
        self._output_scale = float(output_scale)  # sampler outputs are stored as x0 * output_scale (3DHP: 1000)
        self.Spatial_patch_to_embedding = nn.Linear(in_chans + 3, C)
        self.Spatial_pos_embed = nn.Parameter(torch.zeros(1, num_joints, C))
        self.Temporal_pos_embed = nn.Parameter(torch.zeros(1, num_frame, C))
        # indices 1 and 3 carry the two Linear layers, like the reference's Sequential(sinusoid, Linear, GELU, Linear)
        self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(C, C * 2), nn.Identity(), nn.Linear(C * 2, C))
        self.STEblocks = nn.ModuleList([_Block(C, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.TTEblocks = nn.ModuleList([_Block(C, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.Spatial_norm = norm_layer(C)
        self.Temporal_norm = norm_layer(C)
        self.head = nn.Sequential(nn.LayerNorm(C), nn.Linear(C, 3))
        self._engines = {}  # device index -> (Engine, weights fingerprint); shared by DataParallel replicas
        # nn.DataParallel replicas are shallow copies of this module whose parameters are fresh broadcast tensors on
        # every forward (version 0, recycled addresses): they find the module that owns the real parameters through
        # this shared cell and key their packed copies on ITS state.
        self._shared = {"master": self, "epoch": 0}
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.refresh_weights())

    # ------------------------------------------------------------------ engine management
    def refresh_weights(self):
        """Force the packed fp16 copy inside the engine(s) to be rebuilt on the next forward.  Called automatically
        after load_state_dict (also through a parent / nn.DataParallel) and after .to()/.cuda()/.float() (`_apply`);
        optimizer steps and other in-place ops are seen through the parameters' version counters.  Writes that bypass
        autograd's version counter (`p.data.copy_(...)`, `p.data.mul_(...)`, e.g. a hand-written EMA) are invisible to
        any cheap check: call this method after them."""
        self._shared["epoch"] += 1

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if "_shared" in self.__dict__:
            self.refresh_weights()
        return out

    def _named_weights(self):
        """{reference state_dict key: tensor} for this module tree.  Also works inside nn.DataParallel replicas, whose
        parameters are plain attributes recorded in `_former_parameters` (state_dict() / parameters() are empty there)."""
        out = {}
        for mod_name, mod in self.named_modules():
            params = dict(mod._parameters)
            params.update(getattr(mod, "_former_parameters", {}))
            for k, v in params.items():
                if v is not None:
                    out[(mod_name + "." if mod_name else "") + k] = v
        return out

    def _fingerprint(self, weights):
        """Identity of the weights the engine on this device was packed from.  Always the MASTER module's epoch and
        parameter versions (a replica's own tensors say nothing: new every forward); plus, for the master itself,
        the storage addresses (a parameter re-bound with `p.data = new_tensor` keeps its version)."""
        master = self._shared["master"]
        mw = weights if master is self else master._named_weights()
        fp = (self._shared["epoch"],) + tuple(p._version for p in mw.values())
        if master is self:
            fp += tuple(p.data_ptr() for p in mw.values())
        return fp

    def engine(self):
        """The per-device Engine with this module's current weights uploaded (re-packed when any parameter changed)."""
        weights = self._named_weights()
        p0 = weights["Spatial_pos_embed"]
        if p0.device.type != "cuda":
            raise RuntimeError("d3dp_b200.MixSTE2 runs on CUDA (sm_100a) only: move the module with .cuda() first")
        idx = p0.device.index if p0.device.index is not None else torch.cuda.current_device()
        fp = self._fingerprint(weights)
        ent = self._engines.get(idx)
        if ent is None or ent[1] != fp:
            with torch.cuda.device(idx):
                eng = ent[0] if ent is not None else Engine(
                    self.num_frame, self._joints_left, self._joints_right, depth=self.block_depth, scale=self._scale,
                    output_scale=self._output_scale)
                eng.load_pose_estimator_state(weights)
            self._engines[idx] = (eng, fp)
            ent = self._engines[idx]
        return ent[0]

    # ------------------------------------------------------------------ reference surface
    def draw_drop_masks(self, n_streams, device):
        """The DropPath factors of one training-mode forward, drawn like timm's drop_path (per sample of the block
        input's first axis, bernoulli(keep)/keep) in the reference's execution order: for every block i with
        dpr[i] > 0: STEblocks[i] attention [S,F], STEblocks[i] mlp [S,F], TTEblocks[i] attention [S,17],
        TTEblocks[i] mlp [S,17] (common/mixste.py:100,114-115).  Returns the list of depth x 4 tensors."""
        masks = []
        for i, rate in enumerate(self.drop_path_rates):
            keep = 1.0 - rate
            for n in (self.num_frame, self.num_frame, 17, 17):
                if rate == 0.0:
                    masks.append(torch.ones(n_streams, n, device=device))
                else:
                    m = torch.empty(n_streams, n, device=device).bernoulli_(keep)
                    masks.append(m.div_(keep) if keep > 0.0 else m)
        return masks

    def forward(self, x_2d, x_3d, t, drop_masks=None):
        """common/mixste.py:278-298.  eval: x_2d [b,f,17,2], x_3d [b,h,f,17,3], t [b] -> [b,h,f,17,3];
        train layout (is_train=True): x_3d [b,f,17,3] -> [b,f,17,3].  In training mode (`self.training`) with a
        non-zero drop_path_rate the residual branches are dropped per sample exactly like the reference's timm
        DropPath (`drop_masks`: the depth x 4 factor tensors of draw_drop_masks, injected by tests; drawn otherwise).
        Forward only: the kernels keep nothing for a backward pass, so a training-mode call under autograd is refused
        instead of returning a detached tensor."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "d3dp_b200.MixSTE2 is forward-only (sm_100a inference kernels, no backward): call it under "
                "torch.no_grad() / in eval() mode, or train with the reference implementation")
        with torch.no_grad():
            eng = self.engine()
            x_t = x_3d[:, None] if self.is_train else x_3d
            ds = None
            if drop_masks is None and self.training and any(r > 0 for r in self.drop_path_rates):
                drop_masks = self.draw_drop_masks(x_t.shape[0] * x_t.shape[1], eng.device)
            if drop_masks is not None:
                ds = torch.cat([m.to(eng.device, torch.float32).reshape(-1) for m in drop_masks])
            out = eng.denoise(x_2d, x_t, t, drop_scale=ds)
            return out[:, 0] if self.is_train else out
