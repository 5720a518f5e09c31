// Residual + LayerNorm GEMMs (attn.proj and mlp.fc2 of every MixSTE block; reference: common/mixste.py:80,41 with the
// residual adds of Block.forward :114-115 and the LayerNorms that follow, :114,115,243,257,269,273), third design:
// whole 512-wide rows per CTA, weights shared across a CTA pair by cta_group::2 MMAs.
//
// What bounds these kernels is the rate at which an SM can take operand bytes from L2 (~75 GB/s per SM, the same
// figure in the qkv GEMM, in the mainloop-only ablation of the previous LayerNorm kernel and in cuBLAS), because the
// 512 x K weight matrix is re-read for every row tile.  The previous kernel (gemm_ln_pair.cuh) split N across the pair:
// 48 KB of operands per k-block per SM for 128 x 256 outputs.  Here the pair splits M: one UMMA of M = 256 spans both
// SMs (128 rows each), each CTA stages its own 128 rows of A and HALF of each 256-row weight slab, so the same 48 KB
// per k-block per SM now feed 128 x 512 outputs — half the operand traffic per row, and no cross-CTA statistics
// exchange at all, because every CTA owns complete rows (TMEM lane = row, all 512 columns = the two N = 256
// accumulators).  The price: the accumulator fills the SM's 512 TMEM columns, so the MMAs of the next tile start only
// when the epilogue has read the last column of this one; the operand stages and the residual ring prefetch across
// that boundary.
//
//   EPI_RES_LN   x += A.W^T + b ; a16 = fp16(LN_a(x))
//   EPI_RES_LN2  v = x + A.W^T + b ; x = LN_a(v) (+Tpos[f]) ; a16 = fp16(LN_b(x))   (LN_b optional)
//
// Warp 0 = operand TMA producer (both CTAs), warp 1 = MMA issuer (leader CTA) and TMEM owner, warp 2 = residual TMA
// loader, warps 3..10 = epilogue: TMEM lane quadrant = warp % 4, column half g = (warp - 3) / 4, i.e. two threads per
// row with 256 columns each; they combine their (mean, M2) through shared memory with Chan's formula.
#pragma once
#include "gemm_tcgen05.cuh"

namespace d3dp {

template <int STAGES, int RING>
struct LnRowSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;                      // this CTA's 128 rows of A
  static constexpr int B_BYTES = 2 * 128 * 64 * 2;                  // this CTA's half (128 rows) of both 256-row slabs
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;             // 48 KB
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;           // [2 groups] x 16 KB output staging (TMA stores)
  static constexpr int RING_OFFSET = STG_OFFSET + 2 * 16384;        // [2 groups][RING] x 16 KB residual chunks
  static constexpr int SLOT_BYTES = 128 * 32 * 4;
  static constexpr int BAR_OFFSET = RING_OFFSET + 2 * RING * SLOT_BYTES;
  // barriers: full[STAGES] empty[STAGES] tfull tempty rfull[2][RING] rempty[2][RING] ; tmem ptr
  static constexpr int XCH_OFFSET = BAR_OFFSET + 512;               // [2 bufs][2 halves][128 rows] float2
  static constexpr int PARAM_OFFSET = XCH_OFFSET + 2 * 2 * 128 * 8; // 5 x 512 floats: bias, g_a, b_a, g_b, b_b
  static constexpr int TOTAL = PARAM_OFFSET + 5 * 512 * 4 + 1024;
};

template <int EPI, int STAGES, int RING, bool TPOS = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(352, 1)
gemm_ln_row_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                   const GemmParams p) {
  // tmA: A [M,K] fp16 box {64,128}; tmB: W [512,K] fp16 box {64,128}; tmX: x [M,512] fp32 box {32,128} (residual loads
  // and x stores); tmO: a16 [M,512] fp16 box {64,128} (LayerNorm output stores)
  using L = LnRowSmem<STAGES, RING>;
  static_assert(EPI == EPI_RES_LN || EPI == EPI_RES_LN2, "LN epilogues only");
  static_assert(!TPOS || EPI == EPI_RES_LN2, "the temporal position embedding is added by fc2 of block S0 only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                  // per CTA
  uint64_t* tfull_bar = empty_bar + STAGES;                                 // per CTA
  uint64_t* tempty_bar = tfull_bar + 1;                                     // used in the leader only
  uint64_t* rfull_bar = tempty_bar + 1;          // [g*RING + slot]
  uint64_t* rempty_bar = rfull_bar + 2 * RING;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rempty_bar + 2 * RING);
  float2* xch = reinterpret_cast<float2*>(smem + L::XCH_OFFSET);
  float* sprm = reinterpret_cast<float*>(smem + L::PARAM_OFFSET);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (p.M + 127) / 128;
  const int pairs_m = (tiles_m + 1) / 2;
  const int KB = p.K / 64;
  const bool has_b = (EPI == EPI_RES_LN2) && p.ln_b_g != nullptr;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // armed by the leader's producer; bytes arrive from both CTAs
      mbar_init(&empty_bar[s], 1);  // one multicast MMA commit
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * 8);   // the epilogue warps of both CTAs release the pair's accumulator
    for (int i = 0; i < 2 * RING; ++i) {
      mbar_init(&rfull_bar[i], 1);
      mbar_init(&rempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_ptr);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    sprm[i] = p.bias[i];
    sprm[512 + i] = p.ln_a_g[i];
    sprm[1024 + i] = p.ln_a_b[i];
    if (has_b) {
      sprm[1536 + i] = p.ln_b_g[i];
      sprm[2048 + i] = p.ln_b_b[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ operand TMA producer (both CTAs)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int pt = cluster_id; pt < pairs_m; pt += num_clusters) {
        const int row0 = (pt * 2 + rank) * 128;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_backoff(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[s]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * L::STAGE_BYTES);
          tma_load_2d_2sm(sa, &tmA, leader_full, kb * 64, row0);
          tma_load_2d_2sm(sa + L::A_BYTES, &tmB, leader_full, kb * 64, rank * 128);                  // columns [0,256)
          tma_load_2d_2sm(sa + L::A_BYTES + 16384, &tmB, leader_full, kb * 64, 256 + rank * 128);    // columns [256,512)
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(256, 256, 0, 0);
      int s = 0;
      uint32_t ph = 0, tph = 0;
      for (int pt = cluster_id; pt < pairs_m; pt += num_clusters) {
        mbar_wait(tempty_bar, tph ^ 1);  // both CTAs' epilogues have read the previous tile out of TMEM
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
          const uint32_t b_base = a_base + L::A_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = make_sdesc_sw128(a_base + k * 32, 16, 1024);
            mma_f16_ss_2sm(tmem_base, adesc, make_sdesc_sw128(b_base + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
            mma_f16_ss_2sm(tmem_base + 256, adesc, make_sdesc_sw128(b_base + 16384 + k * 32, 16, 1024), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit_2sm_mc(&empty_bar[s], 0x3);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit_2sm_mc(tfull_bar, 0x3);
        tph ^= 1;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual TMA loader (x tile, 32-col chunks)
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int pt = cluster_id; pt < pairs_m; pt += num_clusters) {
        const int row0 = (pt * 2 + rank) * 128;
        for (int c = 0; c < 8; ++c) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int bi = g * RING + slot;
            mbar_wait_backoff(&rempty_bar[bi], ph ^ 1);
            mbar_expect_tx(&rfull_bar[bi], L::SLOT_BYTES);
            tma_load_2d(smem + L::RING_OFFSET + bi * L::SLOT_BYTES, &tmX, &rfull_bar[bi], g * 256 + c * 32, row0);
          }
          if (++slot == RING) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 2 column halves x 4 warps
    const int ew = warp - 3;
    const int g = ew >> 2;            // column half: columns [256g, 256g+256) = accumulator g
    const int quad = warp & 3;        // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;   // tile row == TMEM lane
    const int col0 = g * 256;
    int rslot = 0, xn = 0;
    uint32_t tph = 0, rph = 0;
    // Outputs are staged per group in one 16 KB 128B-swizzled buffer (conflict-free for one-row-per-thread 16 B
    // writes) and leave by TMA store: a 32-column fp32 chunk of x or a 64-column fp16 slab of a16 at a time.
    uint8_t* stg = smem + L::STG_OFFSET + g * 16384 + r * 128;
    const bool leader = (ew & 3) == 0 && lane == 0;
    const uint32_t tempty_leader = mapa_u32(smem_u32(tempty_bar), 0);
    int row0 = 0;
    auto stage_begin = [&]() {  // the previous store of this group has finished reading the buffer
      if (leader) tma_store_wait_read<0>();
      named_bar_sync(2 + g, 128);
    };
    auto stage_x_chunk = [&](const uint32_t (&v)[32], int c) {
      stage_begin();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + ((j ^ (r & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async_smem();
      named_bar_sync(2 + g, 128);
      if (leader) {
        tma_store_2d(&tmX, stg - r * 128, col0 + c * 32, row0);
        tma_store_commit();
      }
    };
    auto stage_a_half = [&](const uint32_t (&o)[16], int c) {  // 32 fp16 columns = half of a 64-column slab
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(stg + ((((c & 1) * 4 + j) ^ (r & 7)) << 4)) =
            make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    };
    auto stage_a_store = [&](int c) {  // after the odd chunk of a slab
      fence_proxy_async_smem();
      named_bar_sync(2 + g, 128);
      if (leader) {
        tma_store_2d(&tmO, stg - r * 128, col0 + (c - 1) * 32, row0);
        tma_store_commit();
      }
    };
    // combine this half-row's (mean, M2) with the other half's (same CTA, the warp 4 further on / back)
    auto exchange = [&](float m_loc, float m2_loc, float eps, float& mean, float& rstd) {
      const int buf = xn & 1;
      xch[(buf * 2 + g) * 128 + r] = make_float2(m_loc, m2_loc);
      named_bar_sync(4 + quad, 64);  // the two warps that share these 32 rows
      const float2 o = xch[(buf * 2 + (g ^ 1)) * 128 + r];
      mean = 0.5f * (m_loc + o.x);
      const float d0 = m_loc - mean, d1 = o.x - mean;
      const float m2 = m2_loc + o.y + 256.0f * (d0 * d0 + d1 * d1);
      rstd = rsqrtf(m2 * (1.0f / 512.0f) + eps);
      ++xn;
    };

    for (int pt = cluster_id; pt < pairs_m; pt += num_clusters) {
      row0 = (pt * 2 + rank) * 128;
      const int grow = row0 + r;
      const bool valid = grow < p.M;
      const bool write_a = p.out16 != nullptr;
      const int f = valid ? (grow % p.F) : 0;
      float rs = 1.0f;  // DropPath scale of this row's branch (x * 1.0f is exact: eval results do not change)
      if (p.row_scale && valid)
        rs = __ldg(p.row_scale + (p.rs_mode == 1 ? (grow / (17 * p.F)) * p.F + f : grow / p.F));
      mbar_wait(tfull_bar, tph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + col0;

      // ---- pass 1: v = (acc + bias) * rs + residual -> TMEM ; shifted one-pass sums (pivot = first value of the half)
      float s1 = 0.f, s2 = 0.f, pv = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        const int bi = g * RING + rslot;
        mbar_wait(&rfull_bar[bi], rph);
        const uint8_t* slot = smem + L::RING_OFFSET + bi * L::SLOT_BYTES + r * 128;
        float res[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(slot + ((j ^ (r & 7)) << 4));
          res[4 * j] = t.x; res[4 * j + 1] = t.y; res[4 * j + 2] = t.z; res[4 * j + 3] = t.w;
        }
        // generic-proxy reads of the slot must be ordered before the TMA (async proxy) refill
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&rempty_bar[bi]);
        if (++rslot == RING) { rslot = 0; rph ^= 1; }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float t = fmaf(__uint_as_float(v[i]) + sprm[col0 + c * 32 + i], rs, res[i]);
          if (c == 0 && i == 0) pv = t;
          const float d = t - pv;
          s1 += d;
          s2 = fmaf(d, d, s2);
          v[i] = __float_as_uint(t);
        }
        tmem_st32(taddr + c * 32, v);
        if constexpr (EPI == EPI_RES_LN) stage_x_chunk(v, c);
      }
      tmem_st_wait();
      const float m_loc = pv + s1 * (1.0f / 256.0f);
      const float m2_loc = fmaxf(s2 - s1 * s1 * (1.0f / 256.0f), 0.f);
      float mean, rstd;
      exchange(m_loc, m2_loc, p.ln_a_eps, mean, rstd);
      // ---- pass 2: y = LN_a(v) ; LN2: shifted sums of y for the second LayerNorm
      float t1 = 0.f, t2 = 0.f, py = 0.f;
      auto ln_a_chunk = [&](uint32_t (&v)[32], int c) {
        const int n = col0 + c * 32;
        uint32_t tp[TPOS ? 32 : 1];
        if constexpr (TPOS) {  // block S0 only: this row's 32 Temporal_pos_embed values as four 256-bit loads (a
          // row-per-thread access: 32 scalar loads per chunk cost 32 L1 wavefronts each, 2.2 ms instead of 1.3 ms)
          const float* src = p.tpos + static_cast<size_t>(f) * 512 + n;
#pragma unroll
          for (int q = 0; q < 4; ++q) ldg256(src + 8 * q, *reinterpret_cast<uint32_t(*)[8]>(tp + 8 * q));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y = (__uint_as_float(v[i]) - mean) * rstd * sprm[512 + n + i] + sprm[1024 + n + i];
          if constexpr (EPI == EPI_RES_LN2) {
            if constexpr (TPOS) y += __uint_as_float(tp[i]);
            if (c == 0 && i == 0) py = y;
            const float d = y - py;
            t1 += d;
            t2 = fmaf(d, d, t2);
          }
          v[i] = __float_as_uint(y);
        }
        if constexpr (EPI == EPI_RES_LN) {
          if (write_a) {
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = pack_half2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            if ((c & 1) == 0) stage_begin();
            stage_a_half(o, c);
            if (c & 1) stage_a_store(c);
          }
        } else {
          if (has_b) tmem_st32(taddr + c * 32, v);
          stage_x_chunk(v, c);
        }
      };
      {  // software-pipelined TMEM reads: chunk c+1 is in flight while chunk c is processed / staged
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < 8; c += 2) {
          tmem_ld_wait();
          tmem_ld32(taddr + (c + 1) * 32, vb);
          ln_a_chunk(va, c);
          tmem_ld_wait();
          if (c + 2 < 8) tmem_ld32(taddr + (c + 2) * 32, va);
          ln_a_chunk(vb, c + 1);
        }
      }
      if constexpr (EPI == EPI_RES_LN2) {
        if (has_b) {  // uniform over the grid
          tmem_st_wait();
          const float m_loc2 = py + t1 * (1.0f / 256.0f);
          const float m2_loc2 = fmaxf(t2 - t1 * t1 * (1.0f / 256.0f), 0.f);
          float mean2, rstd2;
          exchange(m_loc2, m2_loc2, p.ln_b_eps, mean2, rstd2);
          auto ln_b_chunk = [&](const uint32_t (&v)[32], int c) {
            const int n = col0 + c * 32;
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = (__uint_as_float(v[2 * i]) - mean2) * rstd2 * sprm[1536 + n + 2 * i] + sprm[2048 + n + 2 * i];
              const float b = (__uint_as_float(v[2 * i + 1]) - mean2) * rstd2 * sprm[1536 + n + 2 * i + 1] +
                              sprm[2048 + n + 2 * i + 1];
              o[i] = pack_half2(a, b);
            }
            if (write_a) {
              if ((c & 1) == 0) stage_begin();
              stage_a_half(o, c);
              if (c & 1) stage_a_store(c);
            }
          };
          {
            uint32_t va[32], vb[32];
            tmem_ld32(taddr, va);
#pragma unroll 1
            for (int c = 0; c < 8; c += 2) {
              tmem_ld_wait();
              tmem_ld32(taddr + (c + 1) * 32, vb);
              ln_b_chunk(va, c);
              tmem_ld_wait();
              if (c + 2 < 8) tmem_ld32(taddr + (c + 2) * 32, va);
              ln_b_chunk(vb, c + 1);
            }
          }
        }
      }
      // release the accumulator: the leader's MMA thread owns the pair's TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_leader);
      tph ^= 1;
    }
  }

  tma_store_wait_all<0>();  // no-op for threads that issued no bulk stores
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
