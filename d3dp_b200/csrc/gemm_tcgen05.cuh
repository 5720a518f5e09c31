// Persistent TMA + tcgen05 GEMM for the MixSTE Linear layers (reference: common/mixste.py:38,41,65,80 — nn.Linear
// inside Mlp / Attention) with the surrounding elementwise work fused into the epilogue:
//
//   EPI_BIAS_F16       out16 = fp16(A.W^T + b)                                   (qkv Linear, mixste.py:65)
//   EPI_BIAS_GELU_F16  out16 = fp16(gelu_erf(A.W^T + b))                         (fc1 + GELU, mixste.py:38-39)
//   EPI_RES_LN         x += A.W^T + b ; a16 = fp16(LN_a(x))                      (proj + residual, then norm2;
//                                                                                 mixste.py:80,114,115)
//   EPI_RES_LN2        v = x + A.W^T + b ; x = LN_a(v) (+Tpos[f]) ; a16 = fp16(LN_b(x))
//                      (fc2 + residual, shared Spatial_norm/Temporal_norm, optional Temporal_pos_embed,
//                       next block's norm1; mixste.py:41,115,243,250,257,269,273)
//
// A is [M,K] fp16 row-major (K-major), W is the nn.Linear weight [N,K] fp16 row-major (K-major), fp32 accumulate
// in TMEM.  One CTA per SM, 128-row tiles.  Warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2.. = epilogue (TMEM lane == tile row, so LayerNorm is a per-thread loop, no shuffles).
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
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 fp16 = one 128-byte swizzle row

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr, LN exchange [2][2][128] floats
  static constexpr int RED_OFFSET = BAR_OFFSET + 256;
  // per-launch column parameters staged once in smem (uniform LDS.128 broadcasts in the epilogue):
  // [0,2560) bias (N <= 2560) ; LN modes: [512,1024) gamma_a [1024,1536) beta_a [1536,2048) gamma_b [2048,2560) beta_b
  static constexpr int PARAM_OFFSET = RED_OFFSET + 2 * 2 * 128 * 4;
  static constexpr int PARAM_FLOATS = 2560;
  // BN=256 kernels: fp16 output staged per epilogue group in two 128x64 (16 KB, 128B-swizzled) slabs for TMA stores
  static constexpr int OUT_OFFSET = (PARAM_OFFSET + PARAM_FLOATS * 4 + 1023) / 1024 * 1024;
  static constexpr int OUT_BYTES = (BN == 256) ? 2 * 2 * 16384 : 0;
  static constexpr int TOTAL = OUT_OFFSET + OUT_BYTES + 1024 /*alignment slack*/;
};

// erf by Abramowitz & Stegun 7.1.26 (|abs error| <= 1.5e-7): branch-free, 2 MUFU + 7 FMA, so the fc1 epilogue
// keeps up with the MMA pipe.  The result feeds an fp16 store (rel. 4.9e-4), so this is exact-erf GELU for all
// practical purposes (reference: nn.GELU() default = erf form, common/mixste.py:24,39).
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float r = fmaf(-poly, __expf(-ax * ax), 1.0f);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erf_as(v * 0.70710678118654752f)); }

template <int BN, int EPI, int STAGES, int EPI_WARPS>
__global__ void __launch_bounds__(64 + EPI_WARPS * 32, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  using L = GemmSmem<BN, STAGES>;
  constexpr int ACC_STAGES = 512 / BN;
  constexpr int NSPLIT = EPI_WARPS / 4;   // threads sharing one row in the epilogue
  constexpr int COLS_PER_THREAD = BN / NSPLIT;
  static_assert(BN == 256 || BN == 512, "BN");
  static_assert(EPI_WARPS == 4 || EPI_WARPS == 8, "EPI_WARPS");
  static_assert((EPI == EPI_RES_LN || EPI == EPI_RES_LN2) ? BN == 512 : true, "LN epilogues need the full row");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* red = reinterpret_cast<float*>(smem + L::RED_OFFSET);  // [2 (buf)][NSPLIT][128]
  float* sprm = reinterpret_cast<float*>(smem + L::PARAM_OFFSET);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int KB = p.K / GEMM_BK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (BN == 256) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16) {
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) sprm[i] = p.bias[i];
  } else {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
      sprm[i] = p.bias[i];
      sprm[512 + i] = p.ln_a_g[i];
      sprm[1024 + i] = p.ln_a_b[i];
      if (EPI == EPI_RES_LN2 && p.ln_b_g != nullptr) {
        sprm[1536 + i] = p.ln_b_g[i];
        sprm[2048 + i] = p.ln_b_b[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m_blk * GEMM_BM);
#pragma unroll
          for (int nh = 0; nh < BN / 256; ++nh)
            tma_load_2d(sb + nh * 256 * 128, &tmB, &full_bar[s], kb * GEMM_BK, n_blk * BN + nh * 256);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 256, 0, 0);
      int s = 0, as = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t b_base = a_base + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t adesc = make_sdesc_sw128(a_base + k * 32, 16, 1024);
#pragma unroll
            for (int nh = 0; nh < BN / 256; ++nh) {
              const uint64_t bdesc = make_sdesc_sw128(b_base + nh * 256 * 128 + k * 32, 16, 1024);
              mma_f16_ss(d_tmem + nh * 256, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          tc_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&tfull_bar[as]);
        if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 2;
    const int quad = warp & 3;         // TMEM lane quadrant this warp can address
    const int split = ew >> 2;         // which column slice of the row this thread owns
    const int r = quad * 32 + lane;    // tile row == TMEM lane
    const int col0 = split * COLS_PER_THREAD;
    int as = 0;
    uint32_t aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
      const int g = m_blk * GEMM_BM + r;  // global row
      const bool valid = g < p.M;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN + col0;

      if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_GELU_F16) {
        // Each 4-warp group owns COLS_PER_THREAD columns = slabs of 64 columns; a slab is staged in smem in the
        // 128B-swizzled layout (conflict-free for one-row-per-thread 16 B writes) and written out by one TMA store.
        static_assert(COLS_PER_THREAD % 64 == 0, "slab");
        const int n0 = n_blk * BN + col0;
        uint8_t* gbuf = smem + L::OUT_OFFSET + split * 2 * 16384;
        const bool leader = (ew & 3) == 0 && lane == 0;
#pragma unroll 1
        for (int sl = 0; sl < COLS_PER_THREAD / 64; ++sl) {
          uint8_t* buf = gbuf + (sl & 1) * 16384;
          if (leader) tma_store_wait_read<1>();  // the store that last used this buffer has drained
          named_bar_sync(2 + split, 128);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = sl * 2 + cc;
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 bb = *reinterpret_cast<const float2*>(sprm + n0 + c * 32 + 2 * i);
              float a = __uint_as_float(v[2 * i]) + bb.x;
              float b = __uint_as_float(v[2 * i + 1]) + bb.y;
              if constexpr (EPI == EPI_BIAS_GELU_F16) {
                a = gelu_erf(a);
                b = gelu_erf(b);
              }
              o[i] = pack_half2(a, b);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int piece = cc * 4 + i;  // 16-byte piece index inside the 128-byte slab row
              *reinterpret_cast<uint4*>(buf + r * 128 + ((piece ^ (r & 7)) << 4)) =
                  make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(2 + split, 128);
          if (leader) {
            tma_store_2d(&tmC, buf, n0 + sl * 64, m_blk * GEMM_BM);
            tma_store_commit();
          }
        }
      } else {
        // ---- residual + LayerNorm epilogues: this thread owns columns [col0, col0+COLS_PER_THREAD) of row g
        constexpr int NCH = COLS_PER_THREAD / 32;
        float* xrow = p.x + static_cast<size_t>(g) * 512 + col0;
        __half* arow = p.out16 ? p.out16 + static_cast<size_t>(g) * 512 + col0 : nullptr;
        const int f = valid ? (g % p.F) : 0;
        auto row_reduce = [&](float part, int buf) -> float {
          if constexpr (NSPLIT == 1) {
            return part;
          } else {
            float* rb = red + buf * (NSPLIT * 128);
            rb[split * 128 + r] = part;
            named_bar_sync(1, EPI_WARPS * 32);
            float t = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < NSPLIT; ++s2) t += rb[s2 * 128 + r];
            return t;
          }
        };
        // pass 1: v = acc + bias + residual ; keep v in TMEM ; row sum
        float sum = 0.f;
#pragma unroll 2
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          float res[32];
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t t[8];
              ldg256(xrow + c * 32 + 8 * i, t);
#pragma unroll
              for (int q = 0; q < 8; ++q) res[8 * i + q] = __uint_as_float(t[q]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) res[i] = 0.f;
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float t = __uint_as_float(v[i]) + sprm[col0 + c * 32 + i] + res[i];
            sum += t;
            v[i] = __float_as_uint(t);
          }
          tmem_st32(taddr + c * 32, v);
          if constexpr (EPI == EPI_RES_LN) {
            if (valid) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                stg256(xrow + c * 32 + 8 * i, v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4],
                       v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
            }
          }
        }
        tmem_st_wait();
        const float mean = row_reduce(sum, 0) * (1.0f / 512.0f);
        // pass 2: variance about the mean
        float sq = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float d = __uint_as_float(v[i]) - mean;
            sq += d * d;
          }
        }
        const float rstd = rsqrtf(row_reduce(sq, 1) * (1.0f / 512.0f) + p.ln_a_eps);
        // pass 3: y = LN_a(v)
        float sum2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          const int n = col0 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float y = (__uint_as_float(v[i]) - mean) * rstd * sprm[512 + n + i] + sprm[1024 + n + i];
            if constexpr (EPI == EPI_RES_LN2) {
              if (p.tpos) y += __ldg(p.tpos + static_cast<size_t>(f) * 512 + n + i);
              sum2 += y;
            }
            v[i] = __float_as_uint(y);
          }
          if constexpr (EPI == EPI_RES_LN) {
            if (valid) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
                stg256(arow + c * 32 + 16 * i,
                       pack_half2(__uint_as_float(v[16 * i]), __uint_as_float(v[16 * i + 1])),
                       pack_half2(__uint_as_float(v[16 * i + 2]), __uint_as_float(v[16 * i + 3])),
                       pack_half2(__uint_as_float(v[16 * i + 4]), __uint_as_float(v[16 * i + 5])),
                       pack_half2(__uint_as_float(v[16 * i + 6]), __uint_as_float(v[16 * i + 7])),
                       pack_half2(__uint_as_float(v[16 * i + 8]), __uint_as_float(v[16 * i + 9])),
                       pack_half2(__uint_as_float(v[16 * i + 10]), __uint_as_float(v[16 * i + 11])),
                       pack_half2(__uint_as_float(v[16 * i + 12]), __uint_as_float(v[16 * i + 13])),
                       pack_half2(__uint_as_float(v[16 * i + 14]), __uint_as_float(v[16 * i + 15])));
            }
          } else {
            tmem_st32(taddr + c * 32, v);
            if (valid) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                stg256(xrow + c * 32 + 8 * i, v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4],
                       v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
            }
          }
        }
        if constexpr (EPI == EPI_RES_LN2) {
          if (p.ln_b_g != nullptr) {  // uniform across the grid
            tmem_st_wait();
            const float mean2 = row_reduce(sum2, 0) * (1.0f / 512.0f);
            float sq2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
              uint32_t v[32];
              tmem_ld32(taddr + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float d = __uint_as_float(v[i]) - mean2;
                sq2 += d * d;
              }
            }
            const float rstd2 = rsqrtf(row_reduce(sq2, 1) * (1.0f / 512.0f) + p.ln_b_eps);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
              uint32_t v[32];
              tmem_ld32(taddr + c * 32, v);
              tmem_ld_wait();
              const int n = col0 + c * 32;
              uint32_t o[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a = (__uint_as_float(v[2 * i]) - mean2) * rstd2 * sprm[1536 + n + 2 * i] + sprm[2048 + n + 2 * i];
                float b = (__uint_as_float(v[2 * i + 1]) - mean2) * rstd2 * sprm[1536 + n + 2 * i + 1] +
                          sprm[2048 + n + 2 * i + 1];
                o[i] = pack_half2(a, b);
              }
              if (valid) {
#pragma unroll
                for (int i = 0; i < 2; ++i)
                  stg256(arow + c * 32 + 16 * i, o[8 * i], o[8 * i + 1], o[8 * i + 2], o[8 * i + 3], o[8 * i + 4],
                         o[8 * i + 5], o[8 * i + 6], o[8 * i + 7]);
              }
            }
          }
        }
      }
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
  }

  if (BN == 256) tma_store_wait_all<0>();  // no-op for threads that issued no bulk stores
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace d3dp
