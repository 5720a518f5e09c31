// Residual + LayerNorm GEMMs (attn.proj and mlp.fc2 of every MixSTE block; reference: common/mixste.py:80,41 with the
// residual adds of Block.forward :114-115 and the LayerNorms that follow, :114,115,243,257,269,273) as a
// 2-CTA-cluster kernel.
//
// A LayerNorm needs whole 512-wide rows, and a 128x512 fp32 accumulator is all 512 TMEM columns of an SM, which
// would serialise MMA and epilogue.  Here a CTA pair shares each 128-row tile: CTA `rank` computes columns
// [256*rank, 256*rank+256), so each SM holds two 256-column accumulator stages and the epilogue of tile i overlaps
// the MMAs of tile i+1.  Row statistics are combined across the four 128-column quarters (2 epilogue groups x 2
// CTAs) with Chan's parallel mean/M2 formula: every thread owns one row-quarter and publishes its (mean, M2) to both
// CTAs with st.async (data + mbarrier complete_tx in one operation: no fences, no L1 flush); one mbarrier per TMEM
// lane quadrant, so only the eight warps that share rows wait for each other.
// The residual tile is streamed by TMA into a small swizzled ring (no uncoalesced row-per-thread global loads), the
// A tile both CTAs need is fetched once (each CTA loads half of its rows and multicasts them to the pair);
// outputs are staged in swizzled smem and leave by TMA store (row-per-thread 256-bit global stores cost one L1
// wavefront per 32 B sector: ~5 us per tile, measured).  Two operand stages suffice: the kernel is epilogue-bound
// (3 vs 2 stages and L2 prefetch distances 0..16 k-blocks all measure the same), which pays for the staging buffers.  (An L2 prefetch of the next tile's x was
// measured to cost ~0.9 GB of extra DRAM reads per launch — lines evicted before use — and was removed.)
//
//   EPI_RES_LN   x += A.W^T + b ; a16 = fp16(LN_a(x))
//   EPI_RES_LN2  v = x + A.W^T + b ; x = LN_a(v) (+Tpos[f]) ; a16 = fp16(LN_b(x))   (LN_b optional)
#pragma once
#include "gemm_tcgen05.cuh"

namespace d3dp {

template <int ASLOTS, int BSLOTS, int RING>
struct LnPairSmem {
  static constexpr int A_BYTES = 128 * 64 * 2;                      // one k-block of the A tile (16 KB)
  static constexpr int B_BYTES = 256 * 64 * 2;                      // one k-block of this CTA's weight half (32 KB)
  static constexpr int B_OFFSET = ASLOTS * A_BYTES;
  static constexpr int STG_OFFSET = B_OFFSET + BSLOTS * B_BYTES;    // [2 groups] x 16 KB output staging (TMA stores)
  static constexpr int RING_OFFSET = STG_OFFSET + 2 * 16384;        // [2 groups][RING] x 16 KB residual chunks
  static constexpr int SLOT_BYTES = 128 * 32 * 4;
  static constexpr int BAR_OFFSET = RING_OFFSET + 2 * RING * SLOT_BYTES;
  // barriers: afull[ASLOTS] aempty[ASLOTS] bfull[BSLOTS] bempty[BSLOTS] tfull[2] tempty[2] rfull[2][RING]
  // rempty[2][RING] xch[8] ; tmem ptr
  static constexpr int XCH_OFFSET = BAR_OFFSET + 512;               // [2 bufs][4 quarters][128 rows] float2
  static constexpr int PARAM_OFFSET = XCH_OFFSET + 2 * 4 * 128 * 8; // 5 x 256 floats: bias, g_a, b_a, g_b, b_b
  static constexpr int TOTAL = PARAM_OFFSET + 5 * 256 * 4 + 1024;
};

constexpr int LN_PAIR_THREADS = 384;  // 12 warps: A producer, MMA, residual loader, 8 epilogue, B producer

template <int EPI, int ASLOTS, int BSLOTS, int RING>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LN_PAIR_THREADS, 1)
gemm_ln_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                    const GemmParams p) {
  // tmA: A [M,K] fp16 box {64,64}; tmB: W [512,K] fp16 box {64,256}; tmX: x [M,512] fp32 box {32,128} (residual loads
  // and x stores); tmO: a16 [M,512] fp16 box {64,128} (LayerNorm output stores)
  using L = LnPairSmem<ASLOTS, BSLOTS, RING>;
  static_assert(EPI == EPI_RES_LN || EPI == EPI_RES_LN2, "LN epilogues only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* afull_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* aempty_bar = afull_bar + ASLOTS;
  uint64_t* bfull_bar = aempty_bar + ASLOTS;
  uint64_t* bempty_bar = bfull_bar + BSLOTS;
  uint64_t* tfull_bar = bempty_bar + BSLOTS;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;          // [g*RING + slot]
  uint64_t* rempty_bar = rfull_bar + 2 * RING;
  uint64_t* xch_bar = rempty_bar + 2 * RING;     // [buf*4 + quad]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xch_bar + 8);
  float2* xch = reinterpret_cast<float2*>(smem + L::XCH_OFFSET);
  float* sprm = reinterpret_cast<float*>(smem + L::PARAM_OFFSET);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (p.M + 127) / 128;
  const int KB = p.K / 64;
  const int ncol0 = rank * 256;  // first global column of this CTA's half
  const bool has_b = (EPI == EPI_RES_LN2) && p.ln_b_g != nullptr;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < ASLOTS; ++s) {
      mbar_init(&afull_bar[s], 1);
      mbar_init(&aempty_bar[s], 2);  // the A half this CTA multicasts lands in both CTAs: both MMAs must be done
    }
    for (int s = 0; s < BSLOTS; ++s) {
      mbar_init(&bfull_bar[s], 1);
      mbar_init(&bempty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&xch_bar[i], 1);  // armed per exchange; data arrives as tx bytes
    for (int i = 0; i < 2 * RING; ++i) {
      mbar_init(&rfull_bar[i], 1);
      mbar_init(&rempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    sprm[i] = p.bias[ncol0 + i];
    sprm[256 + i] = p.ln_a_g[ncol0 + i];
    sprm[512 + i] = p.ln_a_b[ncol0 + i];
    if (has_b) {
      sprm[768 + i] = p.ln_b_g[ncol0 + i];
      sprm[1024 + i] = p.ln_b_b[ncol0 + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer's barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ operand TMA producer
    // The activation tile A streams from HBM (~2 us under load), the weights from L2: separate rings, so that A can
    // be ASLOTS k-blocks deep (16 KB each) without paying for as many 32 KB weight slots.
    // (warp 0: A, warp 11: B — independent threads, so the A stream runs ahead as far as ITS ring allows.)
    if (lane == 0) {
      int sa = 0;
      uint32_t pha = 0;
      for (int tile = cluster_id; tile < tiles_m; tile += num_clusters) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_backoff(&aempty_bar[sa], pha ^ 1);
          mbar_expect_tx(&afull_bar[sa], L::A_BYTES);
          // the pair shares the A tile: each CTA fetches 64 of its 128 rows and multicasts them to both
          tma_load_2d_mc(smem + sa * L::A_BYTES + rank * (L::A_BYTES / 2), &tmA, &afull_bar[sa], kb * 64,
                         tile * 128 + rank * 64, 0x3);
          if (++sa == ASLOTS) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else if (warp == 11) {
    // ------------------------------------------------------------------ weight TMA producer
    if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      for (int tile = cluster_id; tile < tiles_m; tile += num_clusters) {
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_backoff(&bempty_bar[sb], phb ^ 1);
          mbar_expect_tx(&bfull_bar[sb], L::B_BYTES);
          tma_load_2d(smem + L::B_OFFSET + sb * L::B_BYTES, &tmB, &bfull_bar[sb], kb * 64, ncol0);
          if (++sb == BSLOTS) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 256, 0, 0);
      int sa = 0, sb = 0, as = 0;
      uint32_t pha = 0, phb = 0, aph = 0;
      for (int tile = cluster_id; tile < tiles_m; tile += num_clusters) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&afull_bar[sa], pha);
          mbar_wait(&bfull_bar[sb], phb);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + sa * L::A_BYTES);
          const uint32_t b_base = smem_u32(smem + L::B_OFFSET + sb * L::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_f16_ss(d_tmem, make_sdesc_sw128(a_base + k * 32, 16, 1024), make_sdesc_sw128(b_base + k * 32, 16, 1024),
                       idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit_mc(&aempty_bar[sa], 0x3);
          tc_commit(&bempty_bar[sb]);
          if (++sa == ASLOTS) { sa = 0; pha ^= 1; }
          if (++sb == BSLOTS) { sb = 0; phb ^= 1; }
        }
        tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual TMA loader (x tile, 32-col chunks)
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      for (int tile = cluster_id; tile < tiles_m; tile += num_clusters) {
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int bi = g * RING + slot;
            mbar_wait_backoff(&rempty_bar[bi], ph ^ 1);
            mbar_expect_tx(&rfull_bar[bi], L::SLOT_BYTES);
            tma_load_2d(smem + L::RING_OFFSET + bi * L::SLOT_BYTES, &tmX, &rfull_bar[bi],
                        ncol0 + g * 128 + c * 32, tile * 128);
          }
          if (++slot == RING) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps
    const int ew = warp - 3;
    const int g = ew >> 2;            // column group: local columns [128g, 128g+128)
    const int quad = warp & 3;        // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;   // tile row == TMEM lane
    const int lcol0 = g * 128;        // first local column
    const int qd = rank * 2 + g;      // quarter index of this thread's columns within the 512-wide row
    int as = 0, rslot = 0, xn = 0;
    uint32_t aph = 0, rph = 0;
    // Outputs are staged per group in one 16 KB 128B-swizzled buffer (conflict-free for one-row-per-thread 16 B
    // writes) and leave by TMA store: a 32-column fp32 chunk of x or a 64-column fp16 slab of a16 at a time.
    uint8_t* stg = smem + L::STG_OFFSET + g * 16384 + r * 128;
    const bool leader = (ew & 3) == 0 && lane == 0;
    auto stage_begin = [&]() {  // the previous store of this group has finished reading the buffer
      if (leader) tma_store_wait_read<0>();
      named_bar_sync(2 + g, 128);
    };
    auto stage_x_chunk = [&](const uint32_t (&v)[32], int c, int tile) {
      stage_begin();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(stg + ((j ^ (r & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      fence_proxy_async_smem();
      named_bar_sync(2 + g, 128);
      if (leader) {
        tma_store_2d(&tmX, stg - r * 128, ncol0 + lcol0 + c * 32, tile * 128);
        tma_store_commit();
      }
    };
    auto stage_a_half = [&](const uint32_t (&o)[16], int c) {  // 32 fp16 columns = half of a 64-column slab
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(stg + ((((c & 1) * 4 + j) ^ (r & 7)) << 4)) =
            make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    };
    auto stage_a_store = [&](int c, int tile) {  // after the odd chunk of a slab
      fence_proxy_async_smem();
      named_bar_sync(2 + g, 128);
      if (leader) {
        tma_store_2d(&tmO, stg - r * 128, ncol0 + lcol0 + (c - 1) * 32, tile * 128);
        tma_store_commit();
      }
    };

    // publish this row-quarter's (mean, M2) to both CTAs, wait for the four quarters of the row, return mean / rstd
    auto exchange = [&](float m_loc, float m2_loc, float eps, float& mean, float& rstd) {
      const int buf = xn & 1;
      uint64_t* bar = &xch_bar[buf * 4 + quad];
      const uint32_t off = static_cast<uint32_t>(((buf * 4 + qd) * 128 + r) * 8);
      if (g == 0 && lane == 0) mbar_expect_tx(bar, 2 * 2 * 32 * 8);  // 8 B from each of the 4 warps x 32 lanes of this quadrant
#pragma unroll
      for (uint32_t dst = 0; dst < 2; ++dst)
        st_async_f32x2(mapa_u32(smem_u32(xch), dst) + off, m_loc, m2_loc, mapa_u32(smem_u32(bar), dst));
      mbar_wait(bar, (xn >> 1) & 1);
      const float2* row = xch + buf * 4 * 128 + r;
      float msum = 0.f, m2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 t = row[q * 128];
        msum += t.x;
        m2 += t.y;
      }
      mean = 0.25f * msum;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float dm = row[q * 128].x - mean;
        m2 = fmaf(128.0f * dm, dm, m2);
      }
      rstd = rsqrtf(m2 * (1.0f / 512.0f) + eps);
      ++xn;
    };

    for (int tile = cluster_id; tile < tiles_m; tile += num_clusters) {
      const int grow = tile * 128 + r;
      const bool valid = grow < p.M;
      const bool write_a = p.out16 != nullptr;
      const int f = valid ? (grow % p.F) : 0;
      float rs = 1.0f;  // DropPath scale of this row's branch (x * 1.0f is exact: eval results do not change)
      if (p.row_scale && valid)
        rs = __ldg(p.row_scale + (p.rs_mode == 1 ? (grow / (17 * p.F)) * p.F + f : grow / p.F));
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 256 + lcol0;

      // ---- pass 1: v = acc + bias + residual -> TMEM ; shifted one-pass sums (pivot = first value of the quarter)
      float s1 = 0.f, s2 = 0.f, pv = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
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
          const float t = fmaf(__uint_as_float(v[i]) + sprm[lcol0 + c * 32 + i], rs, res[i]);
          if (c == 0 && i == 0) pv = t;
          const float d = t - pv;
          s1 += d;
          s2 = fmaf(d, d, s2);
          v[i] = __float_as_uint(t);
        }
        tmem_st32(taddr + c * 32, v);
        if constexpr (EPI == EPI_RES_LN) stage_x_chunk(v, c, tile);
      }
      tmem_st_wait();
      const float m_loc = pv + s1 * (1.0f / 128.0f);
      const float m2_loc = fmaxf(s2 - s1 * s1 * (1.0f / 128.0f), 0.f);
      float mean, rstd;
      exchange(m_loc, m2_loc, p.ln_a_eps, mean, rstd);
      // ---- pass 2: y = LN_a(v) ; LN2: shifted sums of y for the second LayerNorm
      float t1 = 0.f, t2 = 0.f, py = 0.f;
      auto ln_a_chunk = [&](uint32_t (&v)[32], int c) {
        const int n = lcol0 + c * 32;
        // Temporal_pos_embed (block S0's fc2 only): the thread's 32 values as four 256-bit loads.  It is a
        // row-per-thread access (every lane another row): 32 scalar loads per chunk touched 32 x 32 L1 wavefronts and
        // made this launch 2.2 ms instead of 1.1 ms; the vector form brings it to 1.3 ms (profiles/r02_notes.md).
        uint32_t tp[32];
        if constexpr (EPI == EPI_RES_LN2) {
          if (p.tpos) {
            const float* src = p.tpos + static_cast<size_t>(f) * 512 + ncol0 + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) ldg256(src + 8 * q, *reinterpret_cast<uint32_t(*)[8]>(tp + 8 * q));
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float y = (__uint_as_float(v[i]) - mean) * rstd * sprm[256 + n + i] + sprm[512 + n + i];
          if constexpr (EPI == EPI_RES_LN2) {
            if (p.tpos) y += __uint_as_float(tp[i]);
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
            if (c & 1) stage_a_store(c, tile);
          }
        } else {
          if (has_b) tmem_st32(taddr + c * 32, v);
          stage_x_chunk(v, c, tile);
        }
      };
      {  // software-pipelined TMEM reads: chunk c+1 is in flight while chunk c is processed / staged
        uint32_t va[32], vb[32];
        tmem_ld32(taddr, va);
#pragma unroll 1
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_wait();
          tmem_ld32(taddr + (c + 1) * 32, vb);
          ln_a_chunk(va, c);
          tmem_ld_wait();
          if (c + 2 < 4) tmem_ld32(taddr + (c + 2) * 32, va);
          ln_a_chunk(vb, c + 1);
        }
      }
      if constexpr (EPI == EPI_RES_LN2) {
        if (has_b) {  // uniform over the grid
          tmem_st_wait();
          const float m_loc2 = py + t1 * (1.0f / 128.0f);
          const float m2_loc2 = fmaxf(t2 - t1 * t1 * (1.0f / 128.0f), 0.f);
          float mean2, rstd2;
          exchange(m_loc2, m2_loc2, p.ln_b_eps, mean2, rstd2);
          auto ln_b_chunk = [&](const uint32_t (&v)[32], int c) {
            const int n = lcol0 + c * 32;
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = (__uint_as_float(v[2 * i]) - mean2) * rstd2 * sprm[768 + n + 2 * i] + sprm[1024 + n + 2 * i];
              const float b = (__uint_as_float(v[2 * i + 1]) - mean2) * rstd2 * sprm[768 + n + 2 * i + 1] +
                              sprm[1024 + n + 2 * i + 1];
              o[i] = pack_half2(a, b);
            }
            if (write_a) {
              if ((c & 1) == 0) stage_begin();
              stage_a_half(o, c);
              if (c & 1) stage_a_store(c, tile);
            }
          };
          {
            uint32_t va[32], vb[32];
            tmem_ld32(taddr, va);
#pragma unroll 1
            for (int c = 0; c < 4; c += 2) {
              tmem_ld_wait();
              tmem_ld32(taddr + (c + 1) * 32, vb);
              ln_b_chunk(va, c);
              tmem_ld_wait();
              if (c + 2 < 4) tmem_ld32(taddr + (c + 2) * 32, va);
              ln_b_chunk(vb, c + 1);
            }
          }
        }
      }
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  }

  tma_store_wait_all<0>();  // no-op for threads that issued no bulk stores
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading / writing this CTA's exchange buffers
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace d3dp
