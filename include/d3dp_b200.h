/*
 * d3dp_b200 — C ABI of the B200-native D3DP diffusion-sampling hot path.
 *
 * The reference (paTRICK-swk/D3DP) is pure Python/PyTorch and has no FFI: its boundary for this path is the Python
 * class `D3DP` (common/diffusionpose.py:55) wrapping the `MixSTE2` denoiser (common/mixste.py:141).  This header is
 * the C boundary that class is re-implemented on (d3dp_b200/diffusionpose.py binds it with ctypes); every entry point
 * names the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name ends in `_host`
 *   - return 0 on success, a negative D3DP_E_* code on failure; nothing throws across the ABI;
 *     d3dp_last_error(h) returns a human-readable message for the last failure on that handle
 *   - all work is enqueued asynchronously on the cudaStream_t given (passed as void*); the caller owns
 *     inputs, outputs and the workspace; the handle owns only its packed copy of the weights
 *   - a handle is bound to the CUDA device current at d3dp_create and is not re-entrant
 *     (one in-flight call per handle); distinct handles are independent
 *   - tensors are dense row-major float32 in the reference's layouts:
 *       x2d  [B, F, 17, 2]    img / x_t / noise [B, H, F, 17, 3]    preds [B, K, H, F, 17, 3]
 */
#ifndef D3DP_B200_H
#define D3DP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D3DP_OK 0
#define D3DP_E_INVALID (-1)     /* bad argument / unsupported shape */
#define D3DP_E_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define D3DP_E_WEIGHTS (-3)     /* unknown weight name, wrong size, or weights incomplete */
#define D3DP_E_WORKSPACE (-4)   /* workspace too small */

typedef struct d3dp_handle d3dp_handle;

/* Mirrors what D3DP.__init__ reads from `args` plus the fixed MixSTE2 hyper-parameters
 * (common/diffusionpose.py:60-88,125-126; common/arguments.py:49,50,58,101,102). */
typedef struct d3dp_config {
  int32_t frames;         /* args.number_of_frames (F), 1..384 in this build                      */
  int32_t joints;         /* 17                                                                   */
  int32_t channels;       /* args.cs, must be 512                                                 */
  int32_t depth;          /* args.dep, 1..8                                                       */
  int32_t heads;          /* 8                                                                    */
  int32_t mlp_hidden;     /* channels * mlp_ratio(2) = 1024                                       */
  int32_t num_timesteps;  /* args.timestep, 1000                                                  */
  float scale;            /* args.scale                                                           */
  int32_t flip_perm[17];  /* joint permutation of the flip TTA: perm[j] = source joint of j       */
  float output_scale;     /* 1.0; 1000.0 reproduces common/diffusionpose_3dhp.py:212,256 (mm units) */
} d3dp_config;

/* D3DP.__init__ / MixSTE2.__init__ (common/diffusionpose.py:60, common/mixste.py:142): allocate a handle.
 * The cosine schedule (common/diffusionpose.py:42-52,75-117) is computed in float64 on creation. */
int d3dp_create(const d3dp_config* cfg, d3dp_handle** out);
void d3dp_destroy(d3dp_handle* h);
const char* d3dp_last_error(const d3dp_handle* h);

/* nn.Module.load_state_dict for `pose_estimator.*` (main.py:630): `name` is the reference state_dict key without the
 * `pose_estimator.` prefix (e.g. "STEblocks.3.attn.qkv.weight"); `data` is float32 with `numel` elements in the
 * reference's layout.  GEMM weights are re-packed to fp16 on the given stream; everything else is copied. */
int d3dp_set_weight(d3dp_handle* h, const char* name, const float* data, int64_t numel, void* stream);
/* number of tensors still missing (0 = ready). */
int d3dp_weights_missing(const d3dp_handle* h);

/* Replace the float64 schedule buffers the sampler reads (registered buffers `alphas_cumprod`,
 * `sqrt_recip_alphas_cumprod`, `sqrt_recipm1_alphas_cumprod`, `sqrt_alphas_cumprod`,
 * `sqrt_one_minus_alphas_cumprod`; common/diffusionpose.py:92-103) — load_state_dict may overwrite them. */
int d3dp_set_schedule(d3dp_handle* h, const double* alphas_cumprod_host, const double* sqrt_recip_host,
                      const double* sqrt_recipm1_host, const double* sqrt_ac_host, const double* sqrt_1mac_host,
                      int32_t n, void* stream);
/* The schedule d3dp_create computes when d3dp_set_schedule is never called: alphas_cumprod[num_timesteps] of
 * cosine_beta_schedule (common/diffusionpose.py:42-52,75-78) in float64.  Host-only (no handle, no GPU), so the C
 * restatement can be pinned against the reference's registered buffer (tests/test_host_cpu.py). */
int d3dp_schedule_host(int32_t num_timesteps, double* alphas_cumprod_out_host);
/* Copy the handle's own float64 alphas_cumprod (host) — used to pin the C schedule against the reference's. */
int d3dp_get_alphas_cumprod(const d3dp_handle* h, double* out_host, int32_t n);

/* Sampling time list: reversed(torch.linspace(-1, T-1, K+1).int()) (common/diffusionpose.py:221-223).
 * Host-only helper; writes K+1 entries. */
int d3dp_time_list(int32_t num_timesteps, int32_t K, int32_t* out_host);

/* Workspace needed by d3dp_denoise / d3dp_ddim_sample for B clips x H hypotheses (flip != 0 doubles the streams). */
int d3dp_workspace_bytes(const d3dp_handle* h, int32_t B, int32_t H, int32_t flip, size_t* bytes);

/* MixSTE2.forward (common/mixste.py:278-298): out[B,H,F,17,3] = D(x2d[B,F,17,2], x_t[B,H,F,17,3], t[B]).
 * t is a device int64 array.  The training layout (is_train=True: x_t [B,F,17,3]) is this call with H = 1.
 * drop_scale: NULL at evaluation (DropPath is Identity, common/diffusionpose.py:121-126).  For a training-mode forward
 * (stochastic depth, timm DropPath at common/mixste.py:100,114-115) it holds the per-sample factors mask/keep_prob of
 * every residual branch, float32, for block d = 0..depth-1 in execution order:
 *   [STEblocks[d] attention: S*F][STEblocks[d] mlp: S*F][TTEblocks[d] attention: S*17][TTEblocks[d] mlp: S*17]
 * with S = B*H streams; the spatial factors are indexed (s*F + f), the temporal ones (s*17 + j) — the first axis of
 * the reference's '(b f) n c' / '(b n) f c' block inputs. */
int d3dp_denoise(d3dp_handle* h, const float* x2d, const float* x_t, const int64_t* t, const float* drop_scale,
                 float* out, int32_t B, int32_t H, void* workspace, size_t workspace_bytes, void* stream);

/* D3DP.ddim_sample_flip (common/diffusionpose.py:215-256) when x2d_flip != NULL, D3DP.ddim_sample (:172-212) when it
 * is NULL.  preds[B,K,H,F,17,3] receives x_start of every step (torch.stack(preds_all, dim=1)).
 *   noise_init  [B,H,F,17,3]      or NULL -> Philox(seed, draw 0)
 *   noise_steps [K-1,B,H,F,17,3]  or NULL -> Philox(seed, draw k+1)
 * Philox noise is addressed by the global hypothesis index h_offset + h inside H_total, so a hypothesis gets the
 * same noise on whatever GPU it is computed (0 <= h_offset, h_offset + H <= H_total).  timesteps_host: K+1
 * descending ints ending in -1, or NULL to use d3dp_time_list.
 * The loop of common/diffusionpose.py:229-254 is captured once per (B, H, K, flip, h_offset, H_total, timesteps,
 * workspace) into a CUDA graph and replayed on `stream`; caller-owned pointers and the seed reach the kernels through
 * a small device-side argument block (written by a one-thread kernel ahead of the graph: the call never blocks
 * the host), so they may change freely between calls.  (Environment D3DP_GRAPH=0 at d3dp_create, or a `stream` that is
 * itself being captured: plain kernel-by-kernel launches, which are then captured into the caller's graph.) */
int d3dp_ddim_sample(d3dp_handle* h, const float* x2d, const float* x2d_flip, const float* noise_init,
                     const float* noise_steps, uint64_t seed, int32_t h_offset, int32_t H_total,
                     const int32_t* timesteps_host, float* preds, int32_t B, int32_t H, int32_t K, void* workspace,
                     size_t workspace_bytes, void* stream);

/* D3DP.q_sample + the clamp/scale of prepare_diffusion_concat (common/diffusionpose.py:260-267,290-306):
 * out = clamp(sqrt(acp_t)*(x0*scale) + sqrt(1-acp_t)*noise, +-1.1*scale)/scale if clamp != 0, else the raw q_sample
 * (no scaling).  x0/noise/out [B, per_sample]; t[B] device int64. */
int d3dp_q_sample(d3dp_handle* h, const float* x0, const float* noise, const int64_t* t, float* out, int32_t B,
                  int64_t per_sample, int32_t clamp, void* stream);

/* JPMA epilogue (main.py:700-712, common/camera.py:30-60, common/loss.py:54-76, main_3dhp.py:782,801-835):
 * zero the root joint, add the trajectory, project with the 9 intrinsics, per-joint argmin over hypotheses of the 2D
 * reprojection error (J-Agg) and the hypothesis mean (P-Agg).
 *   preds [B,K,H,F,17,3]  traj [B,F,3]  cam [B,9]  x2d [B,F,17,2]
 *   jagg_pose, pagg_pose [B,K,F,17,3]   jagg_idx [B,K,F,17] int32   e2d_min [B,K,F,17] or NULL
 * linear != 0 selects project_to_2d_linear (common/camera.py:62-80).
 * hyp_shards = 1: preds is the reference's [B,K,H,F,17,3].  hyp_shards = W > 1: preds is [W,B,K,H/W,F,17,3], the
 * rank-major buffer one NCCL all-gather of the per-rank [B,K,H/W,F,17,3] shards produces (global hypothesis
 * r*(H/W)+hl = shard r, local hl), so the multi-GPU path needs no re-layout copy before the aggregation.
 * The 2-D error and the per-joint argmin are bit-equal to the reference's torch.norm / torch.min on the CPU
 * (first index wins ties). */
int d3dp_jpma(d3dp_handle* h, const float* preds, const float* traj, const float* cam, const float* x2d,
              float* jagg_pose, int32_t* jagg_idx, float* pagg_pose, float* e2d_min, int32_t B, int32_t K, int32_t H,
              int32_t root_joint, int32_t linear, int32_t hyp_shards, void* stream);

/* JPMA with ground truth (evaluation only; main.py:715-718 metrics and main_3dhp.py:785-799 pose export):
 * everything d3dp_jpma produces plus, against gt[B,F,17,3] (root joint zeroed by the caller like main.py:683),
 *   e3d        [B,K,H,F,17]  per-hypothesis 3-D error ||pred - gt|| (input of J-Best / P-Best),
 *   jbest_pose [B,K,F,17,3]  per-joint oracle-best hypothesis (argmin_h e3d, main_3dhp.py:797-799),
 * either may be NULL. */
int d3dp_jpma_gt(d3dp_handle* h, const float* preds, const float* traj, const float* cam, const float* x2d,
                 const float* gt, float* jagg_pose, int32_t* jagg_idx, float* pagg_pose, float* e2d_min, float* e3d,
                 float* jbest_pose, int32_t B, int32_t K, int32_t H, int32_t root_joint, int32_t linear,
                 int32_t hyp_shards, void* stream);

/* Protocol-2 (Procrustes-aligned) per-joint errors (evaluation only; common/loss.py:190-395 p_mpjpe_diffusion_all_min,
 * p_mpjpe_diffusion, p_mpjpe_diffusion_reproj — there a device->numpy round trip with a batched LAPACK SVD):
 * every pose preds[b,k,h,f] (root joint zeroed on read when root_joint >= 0) is rigidly aligned (scale, rotation,
 * translation) to gt[b,f] and perr[B,K,H,F,17] receives the per-joint distance after alignment.  Call with H = 1 on
 * pagg_pose [B,K,F,17,3] for the P-Agg variant (mean_pos=True).  float64 3x3 Jacobi SVD per pose. */
int d3dp_pmpjpe(d3dp_handle* h, const float* preds, const float* gt, float* perr, int32_t B, int32_t K, int32_t H,
                int32_t root_joint, void* stream);

/* Standard-normal Philox fill used for the sampler's noise, exposed so callers/tests can reproduce it:
 * out[B,H,per_bh] for draw index `draw`, hypotheses h_offset..h_offset+H-1 of H_total. */
int d3dp_philox_normal(d3dp_handle* h, float* out, int32_t B, int32_t H, int64_t per_bh, uint64_t seed,
                       int32_t h_offset, int32_t H_total, uint32_t draw, void* stream);

/* Unit-test / profiling hooks for the individual device kernels (same code the sampler launches).
 *   d3dp_test_gemm: out = epilogue(A[M,K] fp16 . W[N,K]^T fp16) with mode
 *     0: out16[M,N] = fp16(acc + bias)            1: out16[M,N] = fp16(gelu(acc + bias))
 *     2: x[M,512] += acc + bias ; out16 = fp16(LN(x; g_a,b_a,eps_a))
 *     3: v = x + acc + bias ; x = LN(v; a) (+tpos[row % F]) ; out16 = fp16(LN(x; b)) (g_b NULL: skipped)
 *   d3dp_test_attn: o16[T,512] = attention(qkv16[T,1536]) ; temporal != 0: sequences of F consecutive rows,
 *     else 17 rows F apart (token order [S,17,F]). */
int d3dp_test_gemm(d3dp_handle* h, int32_t mode, const void* a16, const void* w16, const float* bias, void* out16,
                   float* x, const float* g_a, const float* b_a, float eps_a, const float* g_b, const float* b_b,
                   float eps_b, const float* tpos, int32_t F, int32_t M, int32_t N, int32_t K, void* stream);
int d3dp_test_attn(d3dp_handle* h, int32_t temporal, const void* qkv16, void* o16, int32_t n_streams, void* stream);

/* Library build info: "d3dp_b200 <version> sm_100a". */
const char* d3dp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* D3DP_B200_H */
