// Temporal attention over the F frames of one (stream, joint) sequence, one head per work item
// (reference: common/mixste.py:63-82 Attention.forward as called from the TTEblocks, mixste.py:247-258,270-273).
//
// Token order is [S, J, F] so a temporal sequence is F consecutive rows of the fused QKV activation
// [T, 1536] fp16 (q | k | v thirds, head h = columns 64h..64h+63 of each third, mixste.py:65-67).
// Per work item (sequence, head): TMA loads Q, K, V [ROWS x 64] (ROWS = F rounded up to 16, <= 256) as three
// 128-byte-swizzled tiles; S = Q.K^T goes to TMEM with tcgen05 (M=128 per query tile, N=ROWS, K=64); the softmax
// warpgroup owns one TMEM lane (= query row) per thread; P is written back over S in TMEM as fp16;
// O = P.V is a TS-form tcgen05.mma (A = P from TMEM, B = V from smem, MN-major).
//
// The kernel is bound by the TMEM read port: tcgen05.ld delivers 64 B/clk per SM when the data is consumed
// (profiles/r02_tmem_bw.txt), and a 128 x 256 fp32 score tile is 128 KB = 2 k clk per read.  So every score leaves
// TMEM ONCE: a thread takes its row in two halves of 128 keys, each with its own maximum:
// p_A = 2^((s - m_A) c) -> P_A -> O_A = P_A.V_A, then p_B relative to m = max(m_A, m_B) -> O_B = P_B.V_B, and the
// output is (2^((m_A - m) c) O_A + O_B) / (2^((m_A - m) c) l_A + l_B): exact softmax with 192 KB of TMEM reads per
// query tile (S once + two 32 KB partial outputs) instead of the 288 KB of a max pass + an exp pass (round 1's
// kernel: 0.687 ms; this one 0.636 ms same box, profiles/r02_ab_attnsr.log).  Half B is requested chunk by chunk
// into the registers half A has just vacated, invalid keys are masked with -inf so that each body exists once, the
// four output pieces are requested together, and the TMEM region is freed before the global stores.
// TMEM region of a query tile: S [0,256) -> P_A [0,64) P_B [64,128) O_A [128,192) O_B [192,256).
//
// Schedule (F > 128, two query tiles per item): the single MMA thread issues the two query tiles of an item half a
// period apart — S0(i), PV1(i-1), S1(i), PV0(i) — so warpgroup 0 runs its softmax while the tensor core works for
// warpgroup 1 and vice versa (ping-pong); PV is split into the A and B halves on pa_full / pb_full.
// Q/K and V have their own full/empty barriers per stage: Q and K are released as soon as S1(i) has been issued, V
// after PV1(i), which gives the TMA producer a full item period of prefetch distance with two 96 KB stages.
// Tensor-pipe ceiling of this formulation: the MMAs of an item are 2 048 clk, its TMEM reads 6 144 clk (384 KB),
// its exponentials 4 096 clk of MUFU: <= 33 % tensor pipe however well the three overlap.
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

// barriers: qk_full[2], qk_empty[2], v_full[2], v_empty[2] (per smem stage); s_full[2], pa_full[2], pb_full[2],
// o_full[2], s_free[2] (per query tile / TMEM region)
__global__ void __launch_bounds__(320, 1)
attn_temporal_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnTParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* qkfull_bar = reinterpret_cast<uint64_t*>(smem + ATT_STAGES * ATT_STAGE_BYTES);
  uint64_t* qkempty_bar = qkfull_bar + 2;
  uint64_t* vfull_bar = qkempty_bar + 2;
  uint64_t* vempty_bar = vfull_bar + 2;
  uint64_t* sfull_bar = vempty_bar + 2;
  uint64_t* pafull_bar = sfull_bar + 2;
  uint64_t* pbfull_bar = pafull_bar + 2;
  uint64_t* ofull_bar = pbfull_bar + 2;
  uint64_t* sfree_bar = ofull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sfree_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_items = p.num_seq * 8;
  const int n_mtiles = (p.F + 127) / 128;  // 1 or 2 query tiles
  const int nchunks = (p.rows + 31) / 32;  // 32-key chunks of a score row (1..8)
  const int n_a = nchunks < 4 ? nchunks : 4;  // chunks of half A (keys [0,128)) and half B (keys [128,256))
  const int n_b = nchunks - n_a;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qkfull_bar[i], 1);
      mbar_init(&qkempty_bar[i], 1);
      mbar_init(&vfull_bar[i], 1);
      mbar_init(&vempty_bar[i], 1);
      mbar_init(&sfull_bar[i], 1);
      mbar_init(&pafull_bar[i], 4);
      mbar_init(&pbfull_bar[i], 4);
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
      const int ka = pv_ksteps < 8 ? pv_ksteps : 8;  // k-steps (16 keys each) of half A
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
      // O_A(mt) = P_A.V[0:128) -> columns [128,192), then O_B(mt) = P_B.V[128:rows) -> [192,256); P_A / P_B are fp16
      // pairs at columns [0,64) / [64,128).  `par` is the item parity of the warpgroup's barriers.
      auto issue_pv = [&](int mt, uint32_t stage_base, uint32_t par) {
        const uint32_t p_tmem = tmem_base + mt * 256;
        const uint32_t v_base = stage_base + 2 * ATT_TILE_BYTES;
        mbar_wait(&pafull_bar[mt], par);
        tc_fence_after();
        for (int k = 0; k < ka; ++k)
          mma_f16_ts(p_tmem + 128, p_tmem + k * 8, make_sdesc_sw128(v_base + k * 16 * 128, 1024, 1024), idesc_o,
                     k != 0 ? 1u : 0u);
        if (n_b > 0) {
          mbar_wait(&pbfull_bar[mt], par);
          tc_fence_after();
          for (int k = ka; k < pv_ksteps; ++k)
            mma_f16_ts(p_tmem + 192, p_tmem + k * 8, make_sdesc_sw128(v_base + k * 16 * 128, 1024, 1024), idesc_o,
                       k != ka ? 1u : 0u);
        }
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
          issue_pv(0, base, iph);
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
            issue_pv(1, prev_base, iph ^ 1);
            tc_commit(&vempty_bar[prev_s]);
          }
          // S1(i)
          mbar_wait(&sfree_bar[1], iph ^ 1);
          tc_fence_after();
          issue_s(1, base);
          tc_commit(&qkempty_bar[s]);  // Q and K of this item are not read again
          // PV0(i)
          mbar_wait(&vfull_bar[s], ph);
          issue_pv(0, base, iph);
          first = false;
          prev_base = base;
          prev_s = s;
          if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
          iph ^= 1;
        }
        if (!first) {  // PV1 of the last item
          issue_pv(1, prev_base, iph ^ 1);
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
      const int full_chunks = p.F >> 5;   // chunks whose 32 keys are all valid
      const int tail = p.F & 31;          // valid keys in chunk `full_chunks` (0: none)
      const float c2 = p.scale_log2e;
      uint32_t iph = 0;
      uint32_t v[4][32];                  // one half (up to 128 keys) of this thread's score row

      // keys >= F of the chunk that straddles F become -inf: they drop out of the maximum and exponentiate to 0, so
      // the max / exp bodies below exist once, unmasked (the two-variant bodies did not fit the instruction cache)
      auto mask_chunk = [&](uint32_t (&u)[32], int c) {
        if (c == full_chunks && tail != 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i >= tail) u[i] = 0xff800000u;
        }
      };
      auto max_chunk = [&](const uint32_t (&u)[32], float mx) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(u[i]));
        return mx;
      };
      // p = 2^(s c - moff) for chunk c -> fp16 pairs at TMEM columns [16 c, 16 c + 16); returns the chunk's sum
      auto expo_chunk = [&](const uint32_t (&u)[32], int c, float moff) {
        uint32_t o[16];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = ex2_approx(fmaf(__uint_as_float(u[2 * i]), c2, -moff));
          const float b = ex2_approx(fmaf(__uint_as_float(u[2 * i + 1]), c2, -moff));
          sum += a + b;
          o[i] = pack_half2(a, b);
        }
        tmem_st16(t_s + c * 16, o);
        return sum;
      };

      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int seq = item >> 3, head = item & 7;
        const int qrow = mt * 128 + r;
        mbar_wait(&sfull_bar[mt], iph);
        tc_fence_after();
        // ---- half A: keys [0, 128) into registers, its maximum
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < n_a) tmem_ld32(t_s + j * 32, v[j]);
        tmem_ld_wait();
        float m_a = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < n_a) {
            mask_chunk(v[j], j);
            m_a = max_chunk(v[j], m_a);
          }
        // ---- exponentiate half A chunk by chunk; as soon as a chunk's registers are consumed the matching chunk of
        // half B (keys [128, rows)) is requested into them, so B streams out of TMEM under A's arithmetic
        const float moff_a = m_a * c2;
        float l_a = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < n_a) {
            l_a += expo_chunk(v[j], j, moff_a);
            if (j < n_b) tmem_ld32(t_s + (4 + j) * 32, v[j]);
          }
        float alpha = 1.f, l = l_a;
        // P_A is signalled only once half B is in registers: O_A = P_A.V_A lands on B's TMEM columns
        tmem_ld_wait();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pafull_bar[mt]);
        if (n_b > 0) {
          float m = m_a;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < n_b) {
              mask_chunk(v[j], 4 + j);
              m = max_chunk(v[j], m);
            }
          const float moff = m * c2;
          float l_b = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < n_b) l_b += expo_chunk(v[j], 4 + j, moff);
          alpha = ex2_approx((m_a - m) * c2);
          l = fmaf(alpha, l_a, l_b);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pbfull_bar[mt]);
        }
        // ---- O = (alpha O_A + O_B) / l : all four 32-column pieces requested at once
        mbar_wait(&ofull_bar[mt], iph);
        tc_fence_after();
        tmem_ld32(t_s + 128, v[0]);
        tmem_ld32(t_s + 160, v[1]);
        if (n_b > 0) {
          tmem_ld32(t_s + 192, v[2]);
          tmem_ld32(t_s + 224, v[3]);
        }
        tmem_ld_wait();
        tc_fence_before();  // O is in registers: the region can take the next item's S
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfree_bar[mt]);
        if (qrow < p.F) {
          const float inv = 1.0f / l;
          const float wa = alpha * inv;
          __half* orow = p.out + (static_cast<size_t>(seq) * p.F + qrow) * 512 + head * 64;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              o[i] = __uint_as_float(v[c][i]) * wa;
              if (n_b > 0) o[i] = fmaf(__uint_as_float(v[2 + c][i]), inv, o[i]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float* u = o + 16 * i;
              stg256(orow + c * 32 + 16 * i, pack_half2(u[0], u[1]), pack_half2(u[2], u[3]), pack_half2(u[4], u[5]),
                     pack_half2(u[6], u[7]), pack_half2(u[8], u[9]), pack_half2(u[10], u[11]),
                     pack_half2(u[12], u[13]), pack_half2(u[14], u[15]));
            }
          }
        }
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
