"""Same-box A/B of two BUILDS of libd3dp_b200.so over the six hot kernels at the bench shape (T = 660 960 rows):
CUDA-event time alone, and an integer checksum of every output (builds that only differ in code generation must agree
bit for bit).  Each build runs in its own process, alternating, so clock drift shows up as disagreement between the
two visits of the same build.

    D3DP_NVCC_EXTRA="-DD3DP_SMEM_PTRARITH=1" D3DP_OUT=ab_variant.so bash d3dp_b200/csrc/build.sh
    python profiles/ab_lib.py [libA.so libB.so]        # default: libd3dp_b200.so ab_variant.so (in d3dp_b200/csrc)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def chk(t):
    import torch
    return int(t.contiguous().view(torch.int16 if t.element_size() == 2 else torch.int32).to(torch.int64).sum().item())


def one():
    import torch
    from d3dp_b200 import _lib
    _lib.LIB_PATH = os.environ["AB_LIB"]
    from d3dp_b200.engine import Engine
    eng = Engine(frames=243)
    n_streams = 160
    T = n_streams * 17 * 243
    g = torch.Generator().manual_seed(0)
    a512 = torch.randn(1024, 512, generator=g).half().repeat((T + 1023) // 1024, 1)[:T].cuda()
    a1024 = torch.cat([a512, a512], dim=1)
    qkv = torch.cat([a512, a512, a512], dim=1).contiguous()
    x0 = torch.randn(1024, 512, generator=g).repeat((T + 1023) // 1024, 1)[:T].cuda()
    ga = (1 + 0.1 * torch.randn(512, generator=g)).cuda()
    be = (0.1 * torch.randn(512, generator=g)).cuda()
    tpos = (0.02 * torch.randn(243, 512, generator=g)).cuda()

    def wmat(n, k):
        return (torch.randn(n, k, generator=g) * 0.03).half().cuda()

    jobs = []
    for name, mode, a, w in (("qkv", 0, a512, wmat(1536, 512)), ("fc1_gelu", 1, a512, wmat(1024, 512)),
                             ("proj_res_ln", 2, a512, wmat(512, 512)), ("fc2_res_ln2", 3, a1024, wmat(512, 1024)),
                             ("fc2_tpos", 3, a1024, wmat(512, 1024))):
        bias = (0.1 * torch.randn(w.shape[0], generator=g)).cuda()
        kw = {}
        if mode >= 2:
            kw = dict(ln_a=(ga, be, 1e-6))
        if mode == 3:
            kw.update(ln_b=(be + 1, ga - 1, 1e-6))
        if name == "fc2_tpos":  # the one fc2 launch per forward that adds Temporal_pos_embed (block S0)
            kw.update(tpos=tpos)

        def run(mode=mode, a=a, w=w, bias=bias, kw=kw, keep=False):
            if mode >= 2:
                x = x0.clone() if keep else x0
                out = eng.test_gemm(mode, a, w, bias, x=x, F=243, **kw)
                return (out, x) if keep else None
            out = eng.test_gemm(mode, a, w, bias, F=243)
            return (out,) if keep else None
        jobs.append((name, run))
    for name, temporal in (("attn_temporal", True), ("attn_spatial", False)):
        def run(temporal=temporal, keep=False):
            out = eng.test_attn(temporal, qkv, n_streams)
            return (out,) if keep else None
        jobs.append((name, run))

    only = [k for k in os.environ.get("AB_ONLY", "").split(",") if k]  # e.g. AB_ONLY=proj_res_ln,fc2_res_ln2
    parts = []
    for name, run in jobs:
        if only and name not in only:
            continue
        sums = [chk(t) for t in run(keep=True)]
        for _ in range(3):
            run()
        best = 1e9
        for rnd in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            s.record()
            for _ in range(10):
                run()
            e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e) / 10)
        parts.append(f"{name} {best:.4f} ms chk={sums}")
    if not only or "sampler" in only:
        # whole sampler at a small shape (covers embed / head / time-MLP / DDIM kernels in situ): time per call and the
        # largest difference against the first build visited (variants that re-associate sums are not bit-identical)
        from d3dp_b200 import D3DP
        from d3dp_b200.synthetic import (H36M_JOINTS_LEFT as JL, H36M_JOINTS_RIGHT as JR, flip_2d,
                                         synthetic_pose_estimator_state)
        from d3dp_b200.synthetic import make_args
        F, B, H, K = 243, *[int(v) for v in os.environ.get("AB_SAMPLER", "2,10,2").split(",")]  # AB_SAMPLER="B,H,K"
        model = D3DP(make_args(F), JL, JR, is_train=False, num_proposals=H, sampling_timesteps=K)
        model.pose_estimator.load_state_dict(synthetic_pose_estimator_state(F, seed=0), strict=True)
        model = model.cuda().eval()
        x2d = (0.3 * torch.randn(B, F, 17, 2, generator=torch.Generator().manual_seed(1234)))
        x2f, x2d = flip_2d(x2d).cuda(), x2d.cuda()
        out = model.ddim_sample_flip(x2d, None, input_2d_flip=x2f, seed=7)
        torch.cuda.synchronize()
        ref_path = os.path.join(os.environ.get("AB_TMP", "/tmp"), "ab_lib_sampler_ref.pt")
        if os.path.exists(ref_path):
            diff = (out.cpu() - torch.load(ref_path)).abs().max().item()
        else:
            torch.save(out.cpu(), ref_path)
            diff = 0.0
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(3):
            model.ddim_sample_flip(x2d, None, input_2d_flip=x2f, seed=8 + i)
        e.record()
        torch.cuda.synchronize()
        parts.append(f"sampler(B={B},H={H},K={K}) {s.elapsed_time(e) / 3:.2f} ms max|diff vs first build| {diff:.3e}")
    print(f"{os.path.basename(os.environ['AB_LIB'])}: " + " | ".join(parts), flush=True)


if __name__ == "__main__":
    if os.environ.get("AB_LIB") and len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        csrc = os.path.join(ROOT, "d3dp_b200", "csrc")
        libs = sys.argv[1:3] if len(sys.argv) >= 3 else ["libd3dp_b200.so", "ab_variant.so"]
        libs = [p if os.path.isabs(p) else os.path.join(csrc, p) for p in libs]
        ref = os.path.join(os.environ.get("AB_TMP", "/tmp"), "ab_lib_sampler_ref.pt")
        if os.path.exists(ref):
            os.remove(ref)
        for lib in (libs[0], libs[1]) * int(os.environ.get("AB_VISITS", "2")):
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, AB_LIB=lib),
                           timeout=300)
