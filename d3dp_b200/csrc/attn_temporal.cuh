// Temporal attention over the F frames of one (stream, joint) sequence, one head per work item
// (reference: common/mixste.py:63-82 Attention.forward as called from the TTEblocks, mixste.py:247-258,270-273).
//
// Token order is [S, J, F] so a temporal sequence is F consecutive rows of the fused QKV activation
// [T, 1536] fp16 (q | k | v thirds, head h = columns 64h..64h+63 of each third, mixste.py:65-67).
// Per work item (sequence, head): TMA loads Q, K, V [ROWS x 64] (ROWS = F rounded up to 16, <= 256) as three
// 128-byte-swizzled tiles; S = Q.K^T goes to TMEM with tcgen05 (M=128 per query tile, N=ROWS, K=64); the softmax
// warpgroup owns one TMEM lane (= query row) per thread: max, exp2, row sum, P -> fp16 written back over S in TMEM;
// O = P.V is a TS-form tcgen05.mma (A = P from TMEM, B = V from smem, MN-major); O / rowsum is stored as fp16.
// F <= 256 so one N tile holds the whole row: no online-softmax rescaling.
//
// Schedule (F > 128, two query tiles per item): the single MMA thread issues the two query tiles of an item half a
// period apart — S0(i), PV1(i-1), S1(i), PV0(i) — so warpgroup 0 runs its softmax while the tensor core works for
// warpgroup 1 and vice versa (ping-pong), instead of both warpgroups exponentiating at once and then both waiting
// (same-box A/B at the bench shape: 0.773 -> 0.687 ms).  Measured and rejected on top of this (profiles/r01_notes.md):
// FMA-pipe polynomial exp2 for 40 % of the elements (+1.5 %: not MUFU-bound), packed FFMA2/FADD2 softmax arithmetic
// (+0.7 %: not issue-bound), and a single-TMEM-read two-half online softmax (+12 %: the loads no longer overlap the
// arithmetic inside a warpgroup).  The two passes over S read 576 KB of TMEM per item; at tcgen05.ld's 64 B/clk that is
// 9.2 k clk, which is the measured item time — the TMEM read port is the bound of this kernel.
// Q/K and V have their own full/empty barriers per stage: Q and K are released as soon as S1(i) has been issued, V
// after PV1(i), which gives the TMA producer a full item period of prefetch distance with two 96 KB stages.
#pragma once
#include "ptx.cuh"

namespace d3dp {

struct AttnTParams {
  int num_seq;   // S * J
  int F;         // frames per sequence
  int rows;      // F rounded up to a multiple of 16 (TMA box rows, UMMA N for S, K extent for P.V)
  __half* out;   // [T, 512] fp16
  float scale_log2e;  // head_dim^-0.5 * log2(e)
};

constexpr int ATT_TILE_BYTES = 256 * 128;          // room for 256 rows x 64 fp16
constexpr int ATT_STAGE_BYTES = 3 * ATT_TILE_BYTES;  // Q, K, V
constexpr int ATT_STAGES = 2;
constexpr int ATT_SMEM_BYTES = ATT_STAGES * ATT_STAGE_BYTES + 256 + 1024;

#if defined(D3DP_ATTN_SINGLE_READ) && D3DP_ATTN_SINGLE_READ
#include "attn_temporal_sr.inc"
#else
// barriers: qk_full[2], qk_empty[2], v_full[2], v_empty[2] (per smem stage); s_full[2], p_full[2], o_full[2],
// s_free[2] (per query tile / TMEM region)
__global__ void __launch_bounds__(320, 1)
attn_temporal_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* qkfull_bar = reinterpret_cast<uint64_t*>(smem + ATT_STAGES * ATT_STAGE_BYTES);
  uint64_t* qkempty_bar = qkfull_bar + 2;
  uint64_t* vfull_bar = qkempty_bar + 2;
  uint64_t* vempty_bar = vfull_bar + 2;
  uint64_t* sfull_bar = vempty_bar + 2;
  uint64_t* pfull_bar = sfull_bar + 2;
  uint64_t* ofull_bar = pfull_bar + 2;
  uint64_t* sfree_bar = ofull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.num_seq * 8;
  const int n_mtiles = (p.F + 127) / 128;  // 1 or 2 query tiles

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qkfull_bar[i], 1);
      mbar_init(&qkempty_bar[i], 1);
      mbar_init(&vfull_bar[i], 1);
      mbar_init(&vempty_bar[i], 1);
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&pfull_bar[i], 4);
      mbar_init(&ofull_bar[i], 1);
      mbar_init(&sfree_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tile_bytes = p.rows * 128u;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int seq = item >> 3, head = item & 7;
        uint8_t* st = smem + s * ATT_STAGE_BYTES;
        const int row0 = seq * p.F;
        mbar_wait_backoff(&qkempty_bar[s], ph ^ 1);
        mbar_expect_tx(&qkfull_bar[s], 2u * tile_bytes);
        tma_load_2d(st, &tmQKV, &qkfull_bar[s], head * 64, row0);
        tma_load_2d(st + ATT_TILE_BYTES, &tmQKV, &qkfull_bar[s], 512 + head * 64, row0);
        mbar_wait_backoff(&vempty_bar[s], ph ^ 1);
        mbar_expect_tx(&vfull_bar[s], tile_bytes);
        tma_load_2d(st + 2 * ATT_TILE_BYTES, &tmQKV, &vfull_bar[s], 1024 + head * 64, row0);
        if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, p.rows, 0, 0);  // S = Q.K^T : both K-major
      const uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);      // O = P.V   : A from TMEM, B (V) MN-major
      const int pv_ksteps = p.rows / 16;
      auto issue_s = [&](int mt, uint32_t stage_base) {  // S(mt) = Q[mt].K^T -> TMEM region mt, columns [0, rows)
        const uint32_t d_tmem = tmem_base + mt * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adesc = make_sdesc_sw128(stage_base + mt * 128 * 128 + k * 32, 16, 1024);
          const uint64_t bdesc = make_sdesc_sw128(stage_base + ATT_TILE_BYTES + k * 32, 16, 1024);
          mma_f16_ss(d_tmem, adesc, bdesc, idesc_s, k != 0 ? 1u : 0u);
        }
        tc_commit(&sfull_bar[mt]);
      };
      auto issue_pv = [&](int mt, uint32_t stage_base) {  // O(mt) = P(mt).V : P fp16 at columns [0,128), O at [128,192)
        const uint32_t p_tmem = tmem_base + mt * 256;
        const uint32_t o_tmem = p_tmem + 128;
        const uint32_t v_base = stage_base + 2 * ATT_TILE_BYTES;
        for (int k = 0; k < pv_ksteps; ++k)
          mma_f16_ts(o_tmem, p_tmem + k * 8, make_sdesc_sw128(v_base + k * 16 * 128, 1024, 1024), idesc_o,
                     k != 0 ? 1u : 0u);
        tc_commit(&ofull_bar[mt]);
      };
      int s = 0;
      uint32_t ph = 0, iph = 0;  // ph: phase of the stage barriers; iph: per-item phase of the s/p/o barriers
      if (n_mtiles == 1) {
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
          const uint32_t base = smem_u32(smem + s * ATT_STAGE_BYTES);
          mbar_wait(&qkfull_bar[s], ph);
          mbar_wait(&sfree_bar[0], iph ^ 1);  // previous item's O has been read out
          tc_fence_after();
          issue_s(0, base);
          tc_commit(&qkempty_bar[s]);
          mbar_wait(&vfull_bar[s], ph);
          mbar_wait(&pfull_bar[0], iph);
          tc_fence_after();
          issue_pv(0, base);
          tc_commit(&vempty_bar[s]);
          if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
          iph ^= 1;
        }
      } else {
        bool first = true;
        uint32_t prev_base = 0;
        int prev_s = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
          const uint32_t base = smem_u32(smem + s * ATT_STAGE_BYTES);
          // S0(i)
          mbar_wait(&qkfull_bar[s], ph);
          mbar_wait(&sfree_bar[0], iph ^ 1);
          tc_fence_after();
          issue_s(0, base);
          // PV1(i-1): V of the previous stage (its v_full was waited for before PV0(i-1))
          if (!first) {
            mbar_wait(&pfull_bar[1], iph ^ 1);
            tc_fence_after();
            issue_pv(1, prev_base);
            tc_commit(&vempty_bar[prev_s]);
          }
          // S1(i)
          mbar_wait(&sfree_bar[1], iph ^ 1);
          tc_fence_after();
          issue_s(1, base);
          tc_commit(&qkempty_bar[s]);  // Q and K of this item are not read again
          // PV0(i)
          mbar_wait(&vfull_bar[s], ph);
          mbar_wait(&pfull_bar[0], iph);
          tc_fence_after();
          issue_pv(0, base);
          first = false;
          prev_base = base;
          prev_s = s;
          if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
          iph ^= 1;
        }
        if (!first) {  // PV1 of the last item
          mbar_wait(&pfull_bar[1], iph ^ 1);
          tc_fence_after();
          issue_pv(1, prev_base);
          tc_commit(&vempty_bar[prev_s]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / output warpgroups
    const int mt = (warp - 2) >> 2;  // query tile owned by this warpgroup
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // row within the tile == TMEM lane
    if (mt < n_mtiles) {
      const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + mt * 256;
      const int nchunks = (p.rows + 31) / 32;
      uint32_t iph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int seq = item >> 3, head = item & 7;
        const int qrow = mt * 128 + r;
        mbar_wait(&sfull_bar[mt], iph);
        tc_fence_after();
        // Software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is processed; only the
        // chunk that straddles F carries a key mask.
        const int full_chunks = p.F >> 5;   // chunks whose 32 keys are all valid
        const int tail = p.F & 31;          // valid keys in chunk `full_chunks` (0: none)
        // ---- pass 1: row max over the valid keys
        float mx = -INFINITY;
        {
          uint32_t va[32], vb[32];
          tmem_ld32(t_s, va);
          for (int c = 0; c < nchunks; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nchunks) tmem_ld32(t_s + (c + 1) * 32, vb);
            if (c < full_chunks) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < tail) mx = fmaxf(mx, __uint_as_float(va[i]));
            }
            if (c + 1 < nchunks) {
              tmem_ld_wait();
              if (c + 2 < nchunks) tmem_ld32(t_s + (c + 2) * 32, va);
              if (c + 1 < full_chunks) {
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(vb[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < tail) mx = fmaxf(mx, __uint_as_float(vb[i]));
              }
            }
          }
        }
        const float moff = mx * p.scale_log2e;
        // ---- pass 2: p = exp2(s*c - max*c), row sum, fp16 P written over the consumed part of S
        float sum = 0.f;
        auto expo_chunk = [&](const uint32_t (&v)[32], int c) {
          uint32_t o[16];
          if (c < full_chunks) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = ex2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2e, -moff));
              const float b = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2e, -moff));
              sum += a + b;
              o[i] = pack_half2(a, b);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = 2 * i < tail ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2e, -moff)) : 0.f;
              const float b =
                  2 * i + 1 < tail ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2e, -moff)) : 0.f;
              sum += a + b;
              o[i] = pack_half2(a, b);
            }
          }
          tmem_st16(t_s + c * 16, o);
        };
        {
          uint32_t va[32], vb[32];
          tmem_ld32(t_s, va);
          for (int c = 0; c < nchunks; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nchunks) tmem_ld32(t_s + (c + 1) * 32, vb);
            expo_chunk(va, c);
            if (c + 1 < nchunks) {
              tmem_ld_wait();
              if (c + 2 < nchunks) tmem_ld32(t_s + (c + 2) * 32, va);
              expo_chunk(vb, c + 1);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pfull_bar[mt]);
        // O = P.V done -> normalise and store
        mbar_wait(&ofull_bar[mt], iph);
        tc_fence_after();
        const float inv = 1.0f / sum;
        __half* orow = p.out + (static_cast<size_t>(seq) * p.F + qrow) * 512 + head * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + 128 + c * 32, v);
          tmem_ld_wait();
          if (qrow < p.F) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t* u = v + 16 * i;
              stg256(orow + c * 32 + 16 * i,
                     pack_half2(__uint_as_float(u[0]) * inv, __uint_as_float(u[1]) * inv),
                     pack_half2(__uint_as_float(u[2]) * inv, __uint_as_float(u[3]) * inv),
                     pack_half2(__uint_as_float(u[4]) * inv, __uint_as_float(u[5]) * inv),
                     pack_half2(__uint_as_float(u[6]) * inv, __uint_as_float(u[7]) * inv),
                     pack_half2(__uint_as_float(u[8]) * inv, __uint_as_float(u[9]) * inv),
                     pack_half2(__uint_as_float(u[10]) * inv, __uint_as_float(u[11]) * inv),
                     pack_half2(__uint_as_float(u[12]) * inv, __uint_as_float(u[13]) * inv),
                     pack_half2(__uint_as_float(u[14]) * inv, __uint_as_float(u[15]) * inv));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfree_bar[mt]);
        iph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

#endif

}  // namespace d3dp

// =====================================================================================================================
// Long-sequence variant, 256 < F <= 384 (the sweep's F = 351): one item = (sequence, head) as above, but the score row
// no longer fits one UMMA N tile, so S is produced by two MMAs of N = rows/2 each into adjacent TMEM columns
// [0, rows), P (fp16) overwrites [0, rows/2), O lives at [384, 448).  One shared-memory stage (3 x 48 KB), the up to
// three 128-row query tiles of an item are processed one after the other by a single softmax warpgroup.  Coverage of
// BASELINE config 5, not a tuned kernel.
namespace d3dp {

constexpr int ATTL_TILE_BYTES = 384 * 128;
constexpr int ATTL_SMEM_BYTES = 3 * ATTL_TILE_BYTES + 256 + 1024;

__global__ void __launch_bounds__(192, 1)
attn_temporal_long_kernel(const __grid_constant__ CUtensorMap tmQKV /* box {64, rows/2} */, const AttnTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + 3 * ATTL_TILE_BYTES);
  uint64_t* empty_bar = full_bar + 1;
  uint64_t* sfull_bar = empty_bar + 1;
  uint64_t* pfull_bar = sfull_bar + 1;
  uint64_t* ofull_bar = pfull_bar + 1;
  uint64_t* sfree_bar = ofull_bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.num_seq * 8;
  const int n_mtiles = (p.F + 127) / 128;
  const int half = p.rows / 2;  // rows per TMA box and per S MMA (multiple of 16)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(full_bar, 1);
    mbar_init(empty_bar, 1);
    mbar_init(sfull_bar, 1);
    mbar_init(pfull_bar, 4);
    mbar_init(ofull_bar, 1);
    mbar_init(sfree_bar, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ph = 0;
      const uint32_t bytes = 3u * p.rows * 128u;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int seq = item >> 3, head = item & 7;
        mbar_wait_backoff(empty_bar, ph ^ 1);
        mbar_expect_tx(full_bar, bytes);
        const int row0 = seq * p.F;
        for (int t3 = 0; t3 < 3; ++t3)
          for (int hb = 0; hb < 2; ++hb)
            tma_load_2d(smem + t3 * ATTL_TILE_BYTES + hb * half * 128, &tmQKV, full_bar, t3 * 512 + head * 64,
                        row0 + hb * half);
        ph ^= 1;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_f16(128, half, 0, 0);
      const uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);
      const int pv_ksteps = p.rows / 16;
      uint32_t ph = 0, uph = 0;  // uph: phase of the per-query-tile barriers (one use per query tile)
      const uint32_t q_base = smem_u32(smem), k_base = q_base + ATTL_TILE_BYTES, v_base = q_base + 2 * ATTL_TILE_BYTES;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        mbar_wait(full_bar, ph);
        tc_fence_after();
        for (int mt = 0; mt < n_mtiles; ++mt) {
          mbar_wait(sfree_bar, uph ^ 1);
          tc_fence_after();
          for (int hb = 0; hb < 2; ++hb)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ss(tmem_base + hb * half, make_sdesc_sw128(q_base + mt * 128 * 128 + k * 32, 16, 1024),
                         make_sdesc_sw128(k_base + hb * half * 128 + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          tc_commit(sfull_bar);
          mbar_wait(pfull_bar, uph);
          tc_fence_after();
          for (int k = 0; k < pv_ksteps; ++k)
            mma_f16_ts(tmem_base + 384, tmem_base + k * 8, make_sdesc_sw128(v_base + k * 16 * 128, 1024, 1024), idesc_o,
                       k != 0 ? 1u : 0u);
          tc_commit(ofull_bar);
          uph ^= 1;
        }
        tc_commit(empty_bar);
        ph ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int nchunks = (p.rows + 31) / 32;
    const int full_chunks = p.F >> 5, tail = p.F & 31;
    uint32_t uph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int seq = item >> 3, head = item & 7;
      for (int mt = 0; mt < n_mtiles; ++mt) {
        const int qrow = mt * 128 + r;
        mbar_wait(sfull_bar, uph);
        tc_fence_after();
        float mx = -INFINITY;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + c * 32, v);
          tmem_ld_wait();
          const int nv = c < full_chunks ? 32 : (c == full_chunks ? tail : 0);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nv) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float moff = mx * p.scale_log2e;
        float sum = 0.f;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + c * 32, v);
          tmem_ld_wait();
          const int nv = c < full_chunks ? 32 : (c == full_chunks ? tail : 0);
          uint32_t o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = 2 * i < nv ? ex2_approx(fmaf(__uint_as_float(v[2 * i]), p.scale_log2e, -moff)) : 0.f;
            const float b = 2 * i + 1 < nv ? ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), p.scale_log2e, -moff)) : 0.f;
            sum += a + b;
            o[i] = pack_half2(a, b);
          }
          tmem_st16(t_s + c * 16, o);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pfull_bar);
        mbar_wait(ofull_bar, uph);
        tc_fence_after();
        const float inv = 1.0f / sum;
        __half* orow = p.out + (static_cast<size_t>(seq) * p.F + qrow) * 512 + head * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(t_s + 384 + c * 32, v);
          tmem_ld_wait();
          if (qrow < p.F) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint32_t* u = v + 16 * i;
              stg256(orow + c * 32 + 16 * i, pack_half2(__uint_as_float(u[0]) * inv, __uint_as_float(u[1]) * inv),
                     pack_half2(__uint_as_float(u[2]) * inv, __uint_as_float(u[3]) * inv),
                     pack_half2(__uint_as_float(u[4]) * inv, __uint_as_float(u[5]) * inv),
                     pack_half2(__uint_as_float(u[6]) * inv, __uint_as_float(u[7]) * inv),
                     pack_half2(__uint_as_float(u[8]) * inv, __uint_as_float(u[9]) * inv),
                     pack_half2(__uint_as_float(u[10]) * inv, __uint_as_float(u[11]) * inv),
                     pack_half2(__uint_as_float(u[12]) * inv, __uint_as_float(u[13]) * inv),
                     pack_half2(__uint_as_float(u[14]) * inv, __uint_as_float(u[15]) * inv));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sfree_bar);
        uph ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace d3dp
