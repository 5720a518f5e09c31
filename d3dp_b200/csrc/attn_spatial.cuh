// Spatial attention over the 17 joints of one (stream, frame), all 8 heads per CTA (one warp per head)
// (reference: common/mixste.py:63-82 Attention.forward as called from the STEblocks, mixste.py:239-244,264-269).
//
// 17x17 score tiles are far too small for tcgen05 (M=128 minimum), so this kernel stays on the warp-level
// mma.sync.m16n8k16 path with a shuffle softmax: Q (padded to 32 rows) x K^T (padded to 24 keys), softmax in
// registers, P (32x32) x V (32 keys x 64).  Token order is [S, J, F]: the 17 rows of a spatial sequence are F rows
// apart, each a contiguous 3 KB fp16 qkv row, so the CTA gathers 17 rows with cp.async and every warp works from smem.
// All MMA fragments come from ldmatrix (Q and K plain, V transposed); the padding rows of the 32 x 24 / 32 x 32 tiles
// do not exist in shared memory: their ldmatrix row pointers aim at one 16-byte block of zeros.
// It is an HBM-bound kernel (4 KB/token in+out); the MMA shape padding is irrelevant to its speed.
#pragma once
#include "ptx.cuh"

namespace d3dp {

struct AttnSParams {
  const __half* qkv;  // [T, 1536]
  __half* out;        // [T, 512]
  int num_streams;    // S
  int F;
  float scale;        // head_dim^-0.5
};

constexpr int SP_J = 17;
constexpr int SP_ROW_HALFS = 1536 + 8;  // +16 B pad: consecutive joints land 4 banks apart -> conflict-free fragments
constexpr int SP_BUF_BYTES = SP_J * SP_ROW_HALFS * 2;
constexpr int SP_ZERO_OFFSET = 2 * SP_BUF_BYTES;  // 16 zero bytes: the "row" every padding row of a fragment points to
constexpr int SP_SMEM_BYTES = SP_ZERO_OFFSET + 16;  // double-buffered: the next (stream, frame) streams in during compute

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) attn_spatial_kernel(const AttnSParams p) {
  extern __shared__ __align__(16) uint8_t sp_smem[];
  __half* sm = reinterpret_cast<__half*>(sp_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int num_items = p.num_streams * p.F;

  // gather the 17 qkv rows of one (stream, frame) into buffer `buf`: row(j) = (s*17 + j)*F + f.  A thread copies the
  // same 13 (joint, 16-byte piece) slots of every item: their source / destination offsets are computed once, so a
  // copy costs an add and the cp.async instead of two divisions and 64-bit address arithmetic
  constexpr int SP_COPIES = (SP_J * 192 + 255) / 256;  // 16-byte pieces per thread and item (256 threads)
  uint32_t src_off[SP_COPIES], dst_off[SP_COPIES];     // in halfs / bytes; src relative to the item's (s, j=0, f) row
#pragma unroll
  for (int i = 0; i < SP_COPIES; ++i) {
    const int idx = threadIdx.x + i * 256;
    const int j = idx / 192, ch = idx % 192;  // 192 x 16 B per row
    src_off[i] = static_cast<uint32_t>(j) * static_cast<uint32_t>(p.F) * 1536u + ch * 8;
    dst_off[i] = (j * SP_ROW_HALFS + ch * 8) * 2;
  }
  auto prefetch = [&](int item, int buf) {
    const int s = item / p.F, f = item % p.F;
    const __half* srcb = p.qkv + (static_cast<size_t>(s) * SP_J * p.F + f) * 1536;
    const uint32_t dstb = smem_u32(sm + buf * (SP_BUF_BYTES / 2));
#pragma unroll
    for (int i = 0; i < SP_COPIES; ++i)
      if (i + 1 < SP_COPIES || threadIdx.x + i * 256 < SP_J * 192)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstb + dst_off[i]), "l"(srcb + src_off[i])
                     : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(sp_smem + SP_ZERO_OFFSET)[threadIdx.x] = 0u;
  const uint32_t zero_addr = smem_u32(sp_smem + SP_ZERO_OFFSET);
  int buf = 0;
  if (blockIdx.x < num_items) prefetch(blockIdx.x, 0);
  for (int item = blockIdx.x; item < num_items; item += gridDim.x, buf ^= 1) {
    const int s = item / p.F, f = item % p.F;
    __syncthreads();  // everyone is done with the other buffer (previous item)
    if (item + gridDim.x < num_items) {
      prefetch(item + gridDim.x, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    __half* cur = sm + buf * (SP_BUF_BYTES / 2);

    const int h = warp;
    __half* Q = cur + h * 64;
    const __half* K = cur + 512 + h * 64;
    const __half* V = cur + 1024 + h * 64;
    // shared-memory address of the 16-byte row piece (row, col..col+7) of a [17 x 64] head slice, or the zero block
    auto row_addr = [&](const __half* base, int row, int col) -> uint32_t {
      return row < SP_J ? smem_u32(base + row * SP_ROW_HALFS + col) : zero_addr;
    };

    // ---- S = Q K^T : 2 m-tiles (rows 0-15, 16-31) x 3 n-tiles (keys 0-7, 8-15, 16-23), k = 64 in 4 steps
    float sacc[2][3][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) sacc[mt][nt][i] = 0.f;
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {  // pairs of k-steps: one ldmatrix.x4 of K covers k-steps 2kp and 2kp+1
      uint32_t bk[3][4];              // per n-tile: {b0,b1} of k-step 2kp, {b0,b1} of k-step 2kp+1
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)  // matrices: (k lo, ks 2kp) (k hi, ks 2kp) (k lo, ks 2kp+1) (k hi, ks 2kp+1); rows = keys
        ldmatrix_x4(bk[nt], row_addr(K, nt * 8 + (lane & 7), kp * 32 + (lane >> 3) * 8));
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ks = kp * 2 + kk;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t a[4];  // matrices: (rows 0-7, k lo) (rows 8-15, k lo) (rows 0-7, k hi) (rows 8-15, k hi)
          ldmatrix_x4(a, row_addr(Q, mt * 16 + (lane & 15), ks * 16 + (lane >> 4) * 8));
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) mma_16816(sacc[mt][nt], a, bk[nt][2 * kk], bk[nt][2 * kk + 1]);
        }
      }
    }
    // ---- softmax over the 17 valid keys; C fragment: c0,c1 -> (row g, keys nt*8+2t, +1); c2,c3 -> row g+8
    uint32_t pa[2][2][4];  // P as A fragments: [m-tile][k-step of 16 keys]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int key = nt * 8 + 2 * t + (i & 1);
          if (key < SP_J) mx[i >> 1] = fmaxf(mx[i >> 1], sacc[mt][nt][i]);
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
      }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int key = nt * 8 + 2 * t + (i & 1);
          const float e = key < SP_J ? __expf((sacc[mt][nt][i] - mx[i >> 1]) * p.scale) : 0.f;
          sacc[mt][nt][i] = e;
          sum[i >> 1] += e;
        }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 1);
        sum[hh] += __shfl_xor_sync(0xffffffffu, sum[hh], 2);
        sum[hh] = 1.0f / sum[hh];
      }
      // normalised probabilities -> fp16 A fragments (k-step 0: keys 0-15 = n-tiles 0,1 ; k-step 1: keys 16-31)
      pa[mt][0][0] = pack_half2(sacc[mt][0][0] * sum[0], sacc[mt][0][1] * sum[0]);
      pa[mt][0][1] = pack_half2(sacc[mt][0][2] * sum[1], sacc[mt][0][3] * sum[1]);
      pa[mt][0][2] = pack_half2(sacc[mt][1][0] * sum[0], sacc[mt][1][1] * sum[0]);
      pa[mt][0][3] = pack_half2(sacc[mt][1][2] * sum[1], sacc[mt][1][3] * sum[1]);
      pa[mt][1][0] = pack_half2(sacc[mt][2][0] * sum[0], sacc[mt][2][1] * sum[0]);
      pa[mt][1][1] = pack_half2(sacc[mt][2][2] * sum[1], sacc[mt][2][3] * sum[1]);
      pa[mt][1][2] = 0u;
      pa[mt][1][3] = 0u;
    }
    // ---- O = P V : B fragment b0 = {V[k0+2t][n], V[k0+2t+1][n]}, b1 = same at k0+8 ; n = nt*8 + g
    float oacc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) oacc[mt][nt][i] = 0.f;
    // V fragments with ldmatrix.trans: matrices (keys lo, n-tile 2np) (keys hi, 2np) (keys lo, 2np+1) (keys hi, 2np+1)
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int col = (np * 2 + (lane >> 4)) * 8;
      uint32_t v0[4], v1[4];
      ldmatrix_x4_trans(v0, row_addr(V, (lane & 7) + ((lane >> 3) & 1) * 8, col));       // k-step 0: keys 0..15
      ldmatrix_x4_trans(v1, row_addr(V, 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col));  // k-step 1: key 16 (+ zeros)
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int nt = np * 2 + h2;
        mma_16816(oacc[0][nt], pa[0][0], v0[2 * h2], v0[2 * h2 + 1]);
        mma_16816(oacc[1][nt], pa[1][0], v0[2 * h2], v0[2 * h2 + 1]);
        mma_16816(oacc[0][nt], pa[0][1], v1[2 * h2], v1[2 * h2 + 1]);
        mma_16816(oacc[1][nt], pa[1][1], v1[2 * h2], v1[2 * h2 + 1]);
      }
    }
    // ---- store: stage the head's 17 x 64 outputs in the warp's own (now dead) Q slice, then write 16 B per lane
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(Q + g * SP_ROW_HALFS + col) = pack_half2(oacc[0][nt][0], oacc[0][nt][1]);
      *reinterpret_cast<uint32_t*>(Q + (g + 8) * SP_ROW_HALFS + col) = pack_half2(oacc[0][nt][2], oacc[0][nt][3]);
      if (g == 0) *reinterpret_cast<uint32_t*>(Q + 16 * SP_ROW_HALFS + col) = pack_half2(oacc[1][nt][0], oacc[1][nt][1]);
    }
    __syncwarp();
    for (int idx = lane; idx < SP_J * 8; idx += 32) {
      const int j = idx >> 3, ch = idx & 7;
      const uint4 val = *reinterpret_cast<const uint4*>(Q + j * SP_ROW_HALFS + ch * 8);
      *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(s * SP_J + j) * p.F + f) * 512 + h * 64 + ch * 8) = val;
    }
  }
}

}  // namespace d3dp
