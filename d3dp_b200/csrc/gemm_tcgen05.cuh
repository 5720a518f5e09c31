// Persistent TMA + tcgen05 GEMM for the MixSTE Linear layers whose output is a plain fp16 activation
// (reference: common/mixste.py:65 attn.qkv and :38-39 mlp.fc1 + GELU):
//
//   EPI_BIAS_F16       out16 = fp16(A.W^T + b)
//   EPI_BIAS_GELU_F16  out16 = fp16(gelu_erf(A.W^T + b))
//
// (The two residual + LayerNorm GEMMs, attn.proj and mlp.fc2, live in gemm_ln_pair.cuh.)
//
// A is [M,K] fp16 row-major (K-major), W is the nn.Linear weight [N,K] fp16 row-major (K-major), fp32 accumulate
// in TMEM.  The kernel is a cta_group::2 GEMM: a CTA pair (cluster of 2) computes one 256x256 tile with a single MMA
// stream (UMMA M=256, N=256, K=16).  Each CTA stages its own 128 A rows and HALF of the weight slab (128 of the 256
// N rows) per k-block, so an SM reads 8 KB instead of 12 KB of operands from shared memory per MMA and fills 32 KB
// instead of 48 KB by TMA.  The leader CTA (rank 0) issues all MMAs; both CTAs' TMA loads complete on the leader's
// `full` barriers; MMA commits are multicast to both CTAs; each CTA runs its own epilogue on its 128 rows.
// Warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2..9 = epilogue: TMEM lane == tile row, two 4-warp
// groups of 128 columns each; fp16 results are staged in 128B-swizzled smem slabs and leave by TMA store.
// Template knobs: STAGES operand stages of 32 KB (the mainloop is bound by the bytes TMA keeps in flight: the tile
// needs ~120 GB/s per SM at ~1.5 us of L2/HBM latency, i.e. ~180 KB), OUTBUFS output slabs per column split, and
// BIAS_SMEM (bias staged in shared memory, or read through L1 when the smem is better spent on operand stages).
#pragma once
#include "ptx.cuh"

namespace d3dp {

enum EpiMode : int { EPI_BIAS_F16 = 0, EPI_BIAS_GELU_F16 = 1, EPI_RES_LN = 2, EPI_RES_LN2 = 3 };

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N]
  __half* out16;      // F16 modes: [M, ldo];  LN modes: LN output [M, 512] (may be null -> not written)
  int ldo;
  float* x;  // LN modes: residual stream [M,512] fp32, updated in place
  const void* a_ptr;  // LN modes: the A operand [M,K] fp16 (host side builds a 64-row-box tensor map from it)
  const float* ln_a_g;
  const float* ln_a_b;
  float ln_a_eps;
  const float* ln_b_g;  // EPI_RES_LN2 only; null = no second LN (last block)
  const float* ln_b_b;
  float ln_b_eps;
  const float* tpos;  // EPI_RES_LN2 only; [F,512] added after LN_a, or null
  int F;              // frames (row % F = frame index in the [S, J, F] token order)
  // LN modes, training forward only: stochastic depth (timm DropPath, common/mixste.py:100,114-115).  The branch output
  // acc + bias of row r is multiplied by row_scale[idx(r)] (= mask / keep_prob) before the residual add; null = none.
  // rs_mode 1 (spatial block, one draw per (stream, frame)): idx = (r / (17 F)) * F + r % F;
  // rs_mode 2 (temporal block, one draw per (stream, joint)): idx = r / F.
  const float* row_scale;
  int rs_mode;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;  // 64 fp16 = one 128-byte swizzle row
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + GEMM_EPI_WARPS * 32;

// Exact-erf GELU (reference: nn.GELU() default, common/mixste.py:24,39), two values per call, arranged for the
// epilogue's pipe budget on sm_100 (MUFU: 16 lanes/clk/SM, so every MUFU op costs a warp 8 pipe cycles):
//   gelu(v) = v Phi(v) = max(v, 0) - |v| Phi(-|v|),   Phi(-u) = erfc(u / sqrt 2) / 2 = 2^q(u),
// q = degree-6 fit of log2 Phi(-u) on [0, 5.5] (u clamped there: Phi(-5.5) = 1.9e-8), weighted so that the error of
// the result stays below 0.01 ulp of its fp16 rounding on that range (max |abs error| 1e-6; measured against the fp64
// erf form on 6M normal draws the fp16 results differ in 0.36 % of cases, vs 1.05 % for the Abramowitz-Stegun 7.1.26
// form this replaces).  Cost per value: 1 MUFU (ex2) + 3.5 packed FFMA2 + 2 FMNMX, instead of 2 MUFU + 13 FMA-pipe ops;
// the polynomial runs on negated arguments (w = -|v| = min(v, -v)), hence the alternating coefficient signs.
__device__ __forceinline__ void gelu_erf_x2(float& a, float& b) {
  const float wa = fminf(a, -a), wb = fminf(b, -b);
  const uint64_t w = pack_f32x2(wa, wb);
  const uint64_t u = pack_f32x2(fmaxf(wa, -5.5f), fmaxf(wb, -5.5f));
  uint64_t q = fma_f32x2(u, dup_f32x2(3.298481413e-05f), dup_f32x2(7.550812989e-04f));
  q = fma_f32x2(q, u, dup_f32x2(7.973053230e-03f));
  q = fma_f32x2(q, u, dup_f32x2(5.311827015e-02f));
  q = fma_f32x2(q, u, dup_f32x2(-4.591021916e-01f));
  q = fma_f32x2(q, u, dup_f32x2(1.151066177e+00f));
  q = fma_f32x2(q, u, dup_f32x2(-1.000005425e+00f));
  float qa, qb;
  unpack_f32x2(q, qa, qb);
  const uint64_t g = fma_f32x2(w, pack_f32x2(ex2_approx(qa), ex2_approx(qb)), pack_f32x2(fmaxf(a, 0.f), fmaxf(b, 0.f)));
  unpack_f32x2(g, a, b);
}

// tmA: A [M,K], box {64,128};  tmB: W [N,K], box {64,128} (half a weight slab);  tmC: out [M,N] fp16, box {64,128}

template <int STAGES, int OUTBUFS = 2, bool BIAS_SMEM = true>
struct Gemm2SmSmem {
  static constexpr int A_BYTES = 128 * GEMM_BK * 2;       // this CTA's 128 rows of A
  static constexpr int B_BYTES = 128 * GEMM_BK * 2;       // this CTA's half (128 rows) of the 256-row weight slab
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;  // full[STAGES] empty[STAGES] tfull[2] tempty[2], tmem ptr
  static constexpr int PARAM_OFFSET = BAR_OFFSET + 256;
  static constexpr int PARAM_FLOATS = BIAS_SMEM ? 1536 : 0;  // the whole bias vector (N <= 1536), or read through L1
  static constexpr int OUT_OFFSET = (PARAM_OFFSET + PARAM_FLOATS * 4 + 1023) / 1024 * 1024;
  static constexpr int TOTAL = OUT_OFFSET + 2 * OUTBUFS * 16384 + 1024;  // [2 column splits][OUTBUFS] x 16 KB slabs
};

template <int EPI, int STAGES, int OUTBUFS = 2, bool BIAS_SMEM = true>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using L = Gemm2SmSmem<STAGES, OUTBUFS, BIAS_SMEM>;
  static_assert(EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16, "fp16-output epilogues only");
  constexpr int COLS_PER_THREAD = GEMM_BN / 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                  // per CTA
  uint64_t* tfull_bar = empty_bar + STAGES;                                 // per CTA
  uint64_t* tempty_bar = tfull_bar + 2;                                     // used in the leader only
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sprm = reinterpret_cast<float*>(smem + L::PARAM_OFFSET);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int pairs_m = (tiles_m + 1) / 2;
  const int tiles_n = p.N / GEMM_BN;
  const int num_ctiles = pairs_m * tiles_n;
  const int KB = p.K / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // armed by the leader's producer; bytes arrive from both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast MMA commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * GEMM_EPI_WARPS);  // epilogue warps of both CTAs release the pair's accumulator
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_ptr);
  if constexpr (BIAS_SMEM)
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) sprm[i] = p.bias[i];
  const float* bias_src = BIAS_SMEM ? sprm : p.bias;  // generic loads: ld.shared or ld.global (L1 broadcast)
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        const int m_blk = (ct / tiles_n) * 2 + rank, n_blk = ct % tiles_n;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_backoff(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
          tma_load_2d_2sm(sa, &tmA, leader_full, kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d_2sm(sa + L::A_BYTES, &tmB, leader_full, kb * GEMM_BK, n_blk * GEMM_BN + rank * 128);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, 256, 0, 0);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * GEMM_BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t b_base = a_base + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k)
            mma_f16_ss_2sm(d_tmem, make_sdesc_sw128(a_base + k * 32, 16, 1024),
                           make_sdesc_sw128(b_base + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit_2sm_mc(&empty_bar[s], 0x3);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit_2sm_mc(&tfull_bar[as], 0x3);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int split = ew >> 2;
    const int r = quad * 32 + lane;
    const int col0 = split * COLS_PER_THREAD;
    uint8_t* gbuf = smem + L::OUT_OFFSET + split * OUTBUFS * 16384;
    const bool leader = (ew & 3) == 0 && lane == 0;
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    int as = 0;
    uint32_t aph = 0;
    for (int ct = cluster_id; ct < num_ctiles; ct += num_clusters) {
      const int m_blk = (ct / tiles_n) * 2 + rank, n_blk = ct % tiles_n;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * GEMM_BN + col0;
      const int n0 = n_blk * GEMM_BN + col0;
#pragma unroll 1
      for (int sl = 0; sl < COLS_PER_THREAD / 64; ++sl) {
        uint8_t* buf = gbuf + (sl % OUTBUFS) * 16384;
        // both 32-column chunks of the slab are read from TMEM before any math (one exposed TMEM latency per slab,
        // overlapped with the wait for the staging buffer)
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + sl * 64, v0);
        tmem_ld32(taddr + sl * 64 + 32, v1);
        if (leader) tma_store_wait_read<OUTBUFS - 1>();
        named_bar_sync(2 + split, 128);
        tmem_ld_wait();
        if (sl == COLS_PER_THREAD / 64 - 1) {  // all accumulator columns of this thread are in registers: free the stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(tempty_leader0 + as * 8);  // the leader's MMA thread owns the pair's TMEM
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = sl * 2 + cc;
          const uint32_t(&v)[32] = cc == 0 ? v0 : v1;
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // 8 columns -> one 16-byte piece of the 128-byte slab row
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a, b;
              unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(v[8 * i + 2 * q]), __uint_as_float(v[8 * i + 2 * q + 1])),
                                     *reinterpret_cast<const uint64_t*>(bias_src + n0 + c * 32 + 8 * i + 2 * q)), a, b);
              if constexpr (EPI == EPI_BIAS_GELU_F16) gelu_erf_x2(a, b);
              o[q] = pack_half2(a, b);
            }
            const int piece = cc * 4 + i;
            *reinterpret_cast<uint4*>(buf + r * 128 + ((piece ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2 + split, 128);
        if (leader) {
          tma_store_2d(&tmC, buf, n0 + sl * 64, m_blk * GEMM_BM);
          tma_store_commit();
        }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }

  tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

}  // namespace d3dp
