// CUDA-core kernels around the GEMM/attention pipeline: timestep MLP, input embedding (+flip TTA on the fly),
// regression head, the fused DDIM step, forward noising (q_sample), JPMA aggregation, Philox noise and the
// fp32 -> fp16 weight packer.  Each kernel cites the reference lines it restates.
#pragma once
#include "ptx.cuh"

namespace d3dp {

constexpr int kJ = 17;
constexpr int kC = 512;

// ------------------------------------------------------------------------------------------------ Philox4x32-10
// Counter-based noise so that hypothesis h of clip b gets the same normals whatever GPU computes it
// (counter = element index inside [B, H_total, F, 17, 3] + draw index; key = seed).
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
}
__host__ __device__ __forceinline__ Philox4 philox4x32_10(Philox4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = Philox4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
// One standard normal for (seed, draw, global element index): Box-Muller on two of the four Philox words.
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t draw, uint64_t elem) {
  Philox4 c{static_cast<uint32_t>(elem), static_cast<uint32_t>(elem >> 32), draw, 0u};
  c = philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  const float u1 = (static_cast<float>(c.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  const float u2 = (static_cast<float>(c.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ------------------------------------------------------------------------------------------------ weight packing
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}

// ------------------------------------------------------------------------------------------------ timestep MLP
// tau[b,:] = W2 . gelu(W1 . [sin(t w), cos(t w)] + b1) + b2 , w_i = exp(-i ln(1e4)/255)
// (reference: common/mixste.py:127-139 SinusoidalPositionEmbeddings, :179-184 time_mlp).  One CTA per batch entry.
__global__ void __launch_bounds__(512) time_mlp_kernel(const long long* __restrict__ t, const float* __restrict__ w1,
                                                       const float* __restrict__ b1, const float* __restrict__ w2,
                                                       const float* __restrict__ b2, float* __restrict__ tau) {
  __shared__ float emb[512];
  __shared__ float hid[1024];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float tv = static_cast<float>(t[b]);
  {
    const int i = tid & 255;
    const float w = expf(static_cast<float>(i) * static_cast<float>(-9.210340371976184 / 255.0));
    const float a = tv * w;
    emb[tid] = tid < 256 ? sinf(a) : cosf(a);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  for (int o = warp; o < 1024; o += 16) {  // warp per output: coalesced weight rows
    const float* wr = w1 + static_cast<size_t>(o) * 512;
    float acc = 0.f;
    for (int k = lane; k < 512; k += 32) acc += wr[k] * emb[k];
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) {
      const float v = acc + b1[o];
      hid[o] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    }
  }
  __syncthreads();
  for (int o = warp; o < 512; o += 16) {
    const float* wr = w2 + static_cast<size_t>(o) * 1024;
    float acc = 0.f;
    for (int k = lane; k < 1024; k += 32) acc += wr[k] * hid[k];
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) tau[static_cast<size_t>(b) * 512 + o] = acc + b2[o];
  }
}

// ------------------------------------------------------------------------------------------------ embedding
// x[row,:] = W_e . [x2d(2), x_t(3)] + b_e + Spatial_pos_embed[j] + tau[b]   and   a16 = fp16(LN_norm1(x))
// (reference: common/mixste.py:226-236 eval branch of STE_forward, then STEblocks[0].norm1, :114).
// Streams: s in [0, B*H) read (x2d, clamp(img)/scale); with flip TTA streams [B*H, 2*B*H) read
// (x2d_flip, flip(clamp(img)/scale)) where flip negates coordinate 0 and swaps left/right joints
// (reference: common/diffusionpose.py:148-153).  Token order out: row = (s*17 + j)*F + f.  One warp per token.
// Per-call arguments of the sampler that live in DEVICE memory (one 64-byte block in the workspace, refreshed with a
// single H2D copy before every call): the K-step loop is a CUDA graph replayed across calls, so nothing that changes
// from call to call (caller-owned pointers, the noise seed) may be baked into its kernel parameters.
struct DynArgs {
  const float* x2d;          // [B,F,17,2]
  const float* x2d_flip;     // [B,F,17,2] (flip TTA) or null
  const float* noise_init;   // [B,H,F,17,3] or null -> Philox draw 0
  const float* noise_steps;  // [K-1,B,H,F,17,3] or null -> Philox draw k+1
  float* preds;              // [B,K,H,F,17,3]
  unsigned long long seed;
};

struct EmbedParams {
  const DynArgs* dyn;     // sampler: x2d / x2d_flip are read from here (null: the two fields below are used)
  const float* x2d;       // [B,F,17,2]
  const float* x2d_flip;  // [B,F,17,2] or null
  const float* img;       // [B,H,F,17,3]
  const float* w_e;       // [512,5]
  const float* b_e;       // [512]
  const float* spos;      // [17,512]
  const float* tau;       // [B,512]
  const float* ln_g;      // STEblocks[0].norm1
  const float* ln_b;
  float ln_eps;
  float* x;               // [T,512]
  __half* a16;            // [T,512]
  int B, H, F;
  int n_streams;          // B*H or 2*B*H
  float clamp_hi;         // 1.1*scale  (<=0: no clamp, plain denoise entry)
  float scale;
  int perm[kJ];           // joint permutation of the flip
};

// A lane owns the four channels 4*lane .. 4*lane+3 of each 128-channel quarter of the row, so every store instruction
// of the warp writes one contiguous 512-byte (x, 128-bit per lane) or 256-byte (a16, 64-bit per lane) run: 4 + 4
// fully coalesced stores per token.  (Round 1's mapping channel = i*32 + lane needed 16 + 16 scalar stores; 16
// CONSECUTIVE channels per lane gives 128-bit stores that are 64 bytes apart across lanes, i.e. half-filled sectors.)
// The lane's 16 x 5 weights + bias + LayerNorm affine stay in registers for all its tokens.
__global__ void __launch_bounds__(256) embed_kernel(const EmbedParams p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int BH = p.B * p.H;
  const int T = p.n_streams * kJ * p.F;  // rows: fits 32 bits by far (7 KB of workspace per row); 64-bit div/mod per
                                         // token cost more than the token's arithmetic
  const float* x2d_plain = p.dyn ? p.dyn->x2d : p.x2d;
  const float* x2d_flip = p.dyn ? p.dyn->x2d_flip : p.x2d_flip;
  const int c0 = lane * 4;  // this lane's channels: 128 q + c0 + e, q = 0..3, e = 0..3  (register index i = 4 q + e)
  float we[16][5], be[16], lg[16], lb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = 128 * (i >> 2) + c0 + (i & 3);
#pragma unroll
    for (int k = 0; k < 5; ++k) we[i][k] = p.w_e[c * 5 + k];
    be[i] = p.b_e[c];
    lg[i] = p.ln_g[c];
    lb[i] = p.ln_b[c];
  }
  for (int row = warp; row < T; row += nwarps) {
    const int sj = row / p.F;
    const int f = row - sj * p.F;
    const int s = sj / kJ;
    const int j = sj - s * kJ;
    const bool flip = s >= BH;
    const int bh = flip ? s - BH : s;
    const int b = bh / p.H;
    float in[5];
    {
      const float* src2 = (flip ? x2d_flip : x2d_plain) + ((static_cast<size_t>(b) * p.F + f) * kJ + j) * 2;
      in[0] = src2[0];
      in[1] = src2[1];
      const int js = flip ? p.perm[j] : j;
      const float* src3 = p.img + ((static_cast<size_t>(bh) * p.F + f) * kJ + js) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = src3[c];
        if (p.clamp_hi > 0.f) v = __fdiv_rn(fminf(fmaxf(v, -p.clamp_hi), p.clamp_hi), p.scale);
        in[2 + c] = v;
      }
      if (flip) in[2] = -in[2];
    }
    const float4* sp = reinterpret_cast<const float4*>(p.spos + j * kC + c0);
    const float4* ta = reinterpret_cast<const float4*>(p.tau + static_cast<size_t>(b) * kC + c0);
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 s4 = __ldg(sp + 32 * q), t4 = __ldg(ta + 32 * q);  // quarter q: 128 floats = 32 float4 further on
      const float add[4] = {s4.x + t4.x, s4.y + t4.y, s4.z + t4.z, s4.w + t4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * q + e;
        float acc = be[i];
#pragma unroll
        for (int k = 0; k < 5; ++k) acc += we[i][k] * in[k];
        acc += add[e];
        v[i] = acc;
        sum += acc;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float mean = sum * (1.0f / kC);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sq += (v[i] - mean) * (v[i] - mean);
#pragma unroll
    for (int d = 16; d; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    const float rstd = rsqrtf(sq * (1.0f / kC) + p.ln_eps);
    float4* xr = reinterpret_cast<float4*>(p.x + static_cast<size_t>(row) * kC + c0);
#pragma unroll
    for (int q = 0; q < 4; ++q) xr[32 * q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      h[i] = pack_half2((v[2 * i] - mean) * rstd * lg[2 * i] + lb[2 * i],
                        (v[2 * i + 1] - mean) * rstd * lg[2 * i + 1] + lb[2 * i + 1]);
    uint2* ar = reinterpret_cast<uint2*>(p.a16 + static_cast<size_t>(row) * kC + c0);  // 4 halfs per quarter
#pragma unroll
    for (int q = 0; q < 4; ++q) ar[32 * q] = make_uint2(h[2 * q], h[2 * q + 1]);  // quarter q: 128 halfs = 32 uint2 on
  }
}

// ------------------------------------------------------------------------------------------------ head
// out[s, f, j, :] = W_h . LN(x[row,:]; eps 1e-5) + b_h   (reference: common/mixste.py:207-210,291-296)
// x rows are in [S, J, F] order; out is written in the reference's [.., F, 17, 3] order. One warp per token.
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ x, const float* __restrict__ ln_g,
                                                   const float* __restrict__ ln_b, float ln_eps,
                                                   const float* __restrict__ w_h, const float* __restrict__ b_h,
                                                   float* __restrict__ out, int n_streams, int F) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int T = n_streams * kJ * F;
  // lane owns channels 128 q + 4 lane + e (q, e = 0..3): four coalesced 128-bit loads per row; its slice of the
  // LayerNorm affine and of the 3 x 512 head weights stays in registers for all its rows
  const int c0 = lane * 4;
  float g[16], bt[16], w0[16], w1[16], w2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = 128 * (i >> 2) + c0 + (i & 3);
    g[i] = ln_g[c];
    bt[i] = ln_b[c];
    w0[i] = w_h[c];
    w1[i] = w_h[kC + c];
    w2[i] = w_h[2 * kC + c];
  }
  const float b0 = b_h[0], b1 = b_h[1], b2 = b_h[2];
  for (int row = warp; row < T; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * kC + c0);
    float v[16];
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = __ldg(xr + 32 * q);
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      sum += (t.x + t.y) + (t.z + t.w);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    const float mean = sum * (1.0f / kC);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) sq += (v[i] - mean) * (v[i] - mean);
#pragma unroll
    for (int d = 16; d; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
    const float rstd = rsqrtf(sq * (1.0f / kC) + ln_eps);
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float y = (v[i] - mean) * rstd * g[i] + bt[i];
      o0 += y * w0[i];
      o1 += y * w1[i];
      o2 += y * w2[i];
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      o0 += __shfl_xor_sync(0xffffffffu, o0, d);
      o1 += __shfl_xor_sync(0xffffffffu, o1, d);
      o2 += __shfl_xor_sync(0xffffffffu, o2, d);
    }
    if (lane == 0) {
      const int sj = row / F;
      const int f = row - sj * F;
      const int s = sj / kJ;
      const int j = sj - s * kJ;
      float* o = out + ((static_cast<size_t>(s) * F + f) * kJ + j) * 3;
      o[0] = o0 + b0;
      o[1] = o1 + b1;
      o[2] = o2 + b2;
    }
  }
}

// ------------------------------------------------------------------------------------------------ DDIM step
// One sampler step after the denoiser (reference: common/diffusionpose.py:158-167 un-flip/average/x0/eps and
// :238-254 the eta=1 DDIM update).  den holds the denoiser output for the plain streams [0,BH) and, with TTA,
// the flipped streams [BH,2BH).  eps is formed in float64 from the float64 schedule buffers then rounded to
// float32, exactly like the reference's promoted expression; the three products of the update are rounded
// separately (no FMA contraction) like the reference's eager ops.
struct DdimParams {
  const float* den;     // [n_streams, F, 17, 3]
  float* img;           // [B,H,F,17,3] state, updated in place
  const DynArgs* dyn;   // preds [B,K,H,F,17,3], injected noise_steps (or null -> Philox) and the seed
  int B, H, K, F, k;
  int flip;             // 1: TTA streams present
  int last;             // 1: img = x0
  float scale;
  float out_scale;      // preds are stored as x0 * out_scale (1000 for the 3DHP / mm variant)
  double sqrt_recip_ac, sqrt_recipm1_ac;
  float sqrt_ac_next, c, sigma;
  int h_offset, H_total;  // global hypothesis index = h_offset + h (Philox addressing)
  int perm[kJ];
};

__global__ void __launch_bounds__(256) ddim_step_kernel(const DdimParams p) {
  const long long per_bh = static_cast<long long>(p.F) * kJ * 3;
  const long long n = static_cast<long long>(p.B) * p.H * per_bh;
  const long long BH = static_cast<long long>(p.B) * p.H;
  float* preds = p.dyn->preds;
  const float* noise = (p.dyn->noise_steps && !p.last) ? p.dyn->noise_steps + static_cast<size_t>(p.k) * n : nullptr;
  const unsigned long long seed = p.dyn->seed;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % 3);
    const int j = static_cast<int>((i / 3) % kJ);
    const long long bhf = i / (3 * kJ);
    const int f = static_cast<int>(bhf % p.F);
    const long long bh = bhf / p.F;
    float pred = p.den[i];
    if (p.flip) {
      float o2 = p.den[((BH + bh) * p.F + f) * (kJ * 3) + p.perm[j] * 3 + c];
      if (c == 0) o2 = -o2;
      pred = __fdiv_rn(__fadd_rn(pred, o2), 2.0f);
    }
    float x0 = __fmul_rn(pred, p.scale);
    x0 = fminf(fmaxf(x0, -1.1f * p.scale), 1.1f * p.scale);
    const int b = static_cast<int>(bh / p.H), h = static_cast<int>(bh % p.H);
    preds[(((static_cast<long long>(b) * p.K + p.k) * p.H + h) * per_bh) + (i - bh * per_bh)] =
        p.out_scale == 1.0f ? x0 : __fmul_rn(x0, p.out_scale);
    if (p.last) {
      p.img[i] = x0;
    } else {
      const float xi = p.img[i];
      const float eps = static_cast<float>((p.sqrt_recip_ac * static_cast<double>(xi) - static_cast<double>(x0)) /
                                           p.sqrt_recipm1_ac);
      float z;
      if (noise) {
        z = noise[i];
      } else {
        const unsigned long long ge =
            (static_cast<unsigned long long>(b) * p.H_total + (p.h_offset + h)) * per_bh + (i - bh * per_bh);
        z = philox_normal(seed, static_cast<uint32_t>(p.k + 1), ge);
      }
      p.img[i] = __fadd_rn(__fadd_rn(__fmul_rn(x0, p.sqrt_ac_next), __fmul_rn(p.c, eps)), __fmul_rn(p.sigma, z));
    }
  }
}

// initial state img ~ N(0,I) (reference: common/diffusionpose.py:225) from Philox, draw index 0
__global__ void philox_fill_kernel(float* __restrict__ img, int B, int H, long long per_bh, unsigned long long seed,
                                   int h_offset, int H_total, unsigned draw) {
  const long long n = static_cast<long long>(B) * H * per_bh;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long bh = i / per_bh;
    const int b = static_cast<int>(bh / H), h = static_cast<int>(bh % H);
    const unsigned long long ge =
        (static_cast<unsigned long long>(b) * H_total + (h_offset + h)) * per_bh + (i - bh * per_bh);
    img[i] = philox_normal(seed, draw, ge);
  }
}

// sampler form: injected noise_init if the call gave one, else Philox draw 0 with the call's seed (both via DynArgs)
__global__ void init_img_kernel(float* __restrict__ img, const DynArgs* __restrict__ dyn, int B, int H,
                                long long per_bh, int h_offset, int H_total) {
  const long long n = static_cast<long long>(B) * H * per_bh;
  const float* src = dyn->noise_init;
  const unsigned long long seed = dyn->seed;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (src) {
      img[i] = src[i];
    } else {
      const long long bh = i / per_bh;
      const int b = static_cast<int>(bh / H), h = static_cast<int>(bh % H);
      const unsigned long long ge =
          (static_cast<unsigned long long>(b) * H_total + (h_offset + h)) * per_bh + (i - bh * per_bh);
      img[i] = philox_normal(seed, 0u, ge);
    }
  }
}

// ------------------------------------------------------------------------------------------------ q_sample
// x_t = sqrt(acp_t) * x0 + sqrt(1-acp_t) * noise, optionally clamp(+-1.1 scale)/scale
// (reference: common/diffusionpose.py:260-267 q_sample and :290-306 prepare_diffusion_concat). t per sample.
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const long long* __restrict__ t, const double* __restrict__ sqrt_ac,
                                const double* __restrict__ sqrt_1mac, float* __restrict__ out, int B,
                                long long per_b, float in_scale, float clamp_hi, float out_div) {
  const long long n = static_cast<long long>(B) * per_b;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_b);
    const long long tt = t[b];
    // the reference multiplies float64 buffers into float32 tensors -> float64 arithmetic
    double v = sqrt_ac[tt] * static_cast<double>(__fmul_rn(x0[i], in_scale)) +
               sqrt_1mac[tt] * static_cast<double>(noise[i]);
    if (clamp_hi > 0.f) {
      v = fmin(fmax(v, -static_cast<double>(clamp_hi)), static_cast<double>(clamp_hi));
      v = v / static_cast<double>(out_div);
    }
    out[i] = static_cast<float>(v);
  }
}

// ------------------------------------------------------------------------------------------------ JPMA
// Joint-wise reprojection-based multi-hypothesis aggregation (reference: main.py:700-712 root zero + trajectory +
// project; common/camera.py:44-60 project_to_2d; common/loss.py:54-76 argmin over hypotheses of the 2D error;
// main_3dhp.py:782 P-Agg mean, :801-835 pose-level J-Agg gather).  One thread per (b,k,f,j), loop over H.
struct JpmaParams {
  const float* pred;   // [B,K,H,F,17,3]
  const float* traj;   // [B,F,3]
  const float* cam;    // [B,9]  (f2, c2, k3, p2)
  const float* x2d;    // [B,F,17,2]
  float* jagg_pose;    // [B,K,F,17,3]
  int* jagg_idx;       // [B,K,F,17]
  float* pagg_pose;    // [B,K,F,17,3]
  float* e2d_min;      // [B,K,F,17] or null
  const float* gt;     // [B,F,17,3] or null (evaluation)
  float* e3d;          // [B,K,H,F,17] or null: ||pred - gt|| per hypothesis
  float* jbest_pose;   // [B,K,F,17,3] or null: argmin_h e3d
  int B, K, H, F, root;
  int linear;          // 1: project_to_2d_linear (focal + principal point only)
  int shards;          // pred is [shards, B, K, H/shards, F,17,3] (the rank-major layout an NCCL all-gather of the
                       // per-rank [B,K,h,F,17,3] tensors produces); 1 = the plain [B,K,H,F,17,3] layout
};

__global__ void __launch_bounds__(256) jpma_kernel(const JpmaParams p) {
  const long long n = static_cast<long long>(p.B) * p.K * p.F * kJ;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % kJ);
    const int f = static_cast<int>((i / kJ) % p.F);
    const int k = static_cast<int>((i / (static_cast<long long>(kJ) * p.F)) % p.K);
    const int b = static_cast<int>(i / (static_cast<long long>(kJ) * p.F * p.K));
    const float* tr = p.traj + (static_cast<size_t>(b) * p.F + f) * 3;
    const float* cm = p.cam + static_cast<size_t>(b) * 9;
    const float u = p.x2d[((static_cast<size_t>(b) * p.F + f) * kJ + j) * 2];
    const float v = p.x2d[((static_cast<size_t>(b) * p.F + f) * kJ + j) * 2 + 1];
    float best = INFINITY;
    int best_h = 0;
    float bx = 0.f, by = 0.f, bz = 0.f;
    float best3 = INFINITY, jx = 0.f, jy = 0.f, jz = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    if (p.gt) {
      const float* gp = p.gt + ((static_cast<size_t>(b) * p.F + f) * kJ + j) * 3;
      gx = gp[0]; gy = gp[1]; gz = gp[2];
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    const int h_loc = p.H / p.shards;
    for (int h = 0; h < p.H; ++h) {
      const int r = h / h_loc, hl = h - r * h_loc;  // global hypothesis h = shard r, local index hl
      const float* q =
          p.pred + (((((static_cast<size_t>(r) * p.B + b) * p.K + k) * h_loc + hl) * p.F + f) * kJ + j) * 3;
      float px = q[0], py = q[1], pz = q[2];
      if (j == p.root) px = py = pz = 0.f;
      sx = __fadd_rn(sx, px);
      sy = __fadd_rn(sy, py);
      sz = __fadd_rn(sz, pz);
      const float X = __fadd_rn(px, tr[0]), Y = __fadd_rn(py, tr[1]), Z = __fadd_rn(pz, tr[2]);
      const float xx = fminf(fmaxf(__fdiv_rn(X, Z), -1.f), 1.f);
      const float yy = fminf(fmaxf(__fdiv_rn(Y, Z), -1.f), 1.f);
      float pu, pv;
      if (p.linear) {
        pu = __fadd_rn(__fmul_rn(cm[0], xx), cm[2]);
        pv = __fadd_rn(__fmul_rn(cm[1], yy), cm[3]);
      } else {
        const float r2 = __fadd_rn(__fmul_rn(xx, xx), __fmul_rn(yy, yy));
        const float r4 = __fmul_rn(r2, r2), r6 = __fmul_rn(__fmul_rn(r2, r2), r2);
        const float radial =
            __fadd_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(cm[4], r2), __fmul_rn(cm[5], r4)), __fmul_rn(cm[6], r6)));
        const float tan = __fadd_rn(__fmul_rn(cm[7], xx), __fmul_rn(cm[8], yy));
        const float rt = __fadd_rn(radial, tan);
        const float X3 = __fadd_rn(__fmul_rn(xx, rt), __fmul_rn(cm[7], r2));
        const float Y3 = __fadd_rn(__fmul_rn(yy, rt), __fmul_rn(cm[8], r2));
        pu = __fadd_rn(__fmul_rn(cm[0], X3), cm[2]);
        pv = __fadd_rn(__fmul_rn(cm[1], Y3), cm[3]);
      }
      const float du = __fsub_rn(pu, u), dv = __fsub_rn(pv, v);
      // torch.norm(dim=-1) accumulates acc = fma(x, x, acc) in element order (checked bit for bit against ATen's CPU
      // kernel on 110 160 values, tests/test_oracle_cpu.py): e = sqrt(fma(dv, dv, du*du))
      const float e = __fsqrt_rn(__fmaf_rn(dv, dv, __fmul_rn(du, du)));
      if (p.gt) {
        const float ax = __fsub_rn(px, gx), ay = __fsub_rn(py, gy), az = __fsub_rn(pz, gz);
        const float e3 = __fsqrt_rn(__fmaf_rn(az, az, __fmaf_rn(ay, ay, __fmul_rn(ax, ax))));
        if (p.e3d) p.e3d[(((static_cast<size_t>(b) * p.K + k) * p.H + h) * p.F + f) * kJ + j] = e3;
        if (e3 < best3) { best3 = e3; jx = px; jy = py; jz = pz; }
      }
      if (e < best) {  // strict: first index wins ties, like torch.min
        best = e;
        best_h = h;
        bx = px; by = py; bz = pz;
      }
    }
    float* jo = p.jagg_pose + i * 3;
    jo[0] = bx; jo[1] = by; jo[2] = bz;
    p.jagg_idx[i] = best_h;
    const float invH = 1.0f / static_cast<float>(p.H);
    float* po = p.pagg_pose + i * 3;
    po[0] = sx * invH; po[1] = sy * invH; po[2] = sz * invH;
    if (p.e2d_min) p.e2d_min[i] = best;
    if (p.gt && p.jbest_pose) {
      float* bo = p.jbest_pose + i * 3;
      bo[0] = jx; bo[1] = jy; bo[2] = jz;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Protocol-2 errors: per-pose rigid alignment (scale, rotation, translation) of every hypothesis pose to the ground
// truth, then the per-joint distance (reference: common/loss.py:190-259 p_mpjpe_diffusion_all_min, :261-330
// p_mpjpe_diffusion, :332-395 p_mpjpe_diffusion_reproj — there a GPU -> numpy round trip with a batched LAPACK SVD).
// One thread per pose (b,k,h,f).  With X = target, Y = prediction, X0/Y0 centred and Frobenius-normalised,
// M = X0^T Y0 = U S V^T:  R = V diag(1,1,d) U^T with d = sign det(V U^T) = sign det(M),  a = (s1+s2+d s3) |X0|/|Y0|,
// t = muX - a muY R,  aligned = a Y R + t.  The 3x3 SVD is a cyclic Jacobi eigen-decomposition of M^T M in float64
// (V, s^2), u_i = M v_i / s_i for the two leading directions and u3 = u1 x u2 (the sign of u3 cancels in R).
struct ProcrustesParams {
  const float* pred;  // [B,K,H,F,17,3]
  const float* gt;    // [B,F,17,3]
  float* err;         // [B,K,H,F,17]
  int B, K, H, F, root;
};

__device__ __forceinline__ void jacobi_rotate(double (&A)[3][3], double (&V)[3][3], int p, int q) {
  if (fabs(A[p][q]) < 1e-300) return;
  const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
  const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
  const int r = 3 - p - q;
  const double app = A[p][p], aqq = A[q][q], apq = A[p][q], arp = A[r][p], arq = A[r][q];
  A[p][p] = app - t * apq;
  A[q][q] = aqq + t * apq;
  A[p][q] = A[q][p] = 0.0;
  A[r][p] = A[p][r] = c * arp - s * arq;
  A[r][q] = A[q][r] = s * arp + c * arq;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double vp = V[i][p], vq = V[i][q];
    V[i][p] = c * vp - s * vq;
    V[i][q] = s * vp + c * vq;
  }
}

__global__ void __launch_bounds__(128) procrustes_kernel(const ProcrustesParams p) {
  const long long n = static_cast<long long>(p.B) * p.K * p.H * p.F;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i % p.F);
    const int b = static_cast<int>(i / (static_cast<long long>(p.F) * p.H * p.K));
    const float* Yp = p.pred + i * (kJ * 3);
    const float* Xp = p.gt + (static_cast<size_t>(b) * p.F + f) * (kJ * 3);
    float Y[kJ][3], X[kJ][3];
    double muX[3] = {0, 0, 0}, muY[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < kJ; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Y[j][c] = j == p.root ? 0.f : Yp[j * 3 + c];
        X[j][c] = Xp[j * 3 + c];
        muX[c] += X[j][c];
        muY[c] += Y[j][c];
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) { muX[c] /= kJ; muY[c] /= kJ; }
    double nX = 0, nY = 0, M[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      double x0[3], y0[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        x0[c] = X[j][c] - muX[c];
        y0[c] = Y[j][c] - muY[c];
        nX += x0[c] * x0[c];
        nY += y0[c] * y0[c];
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[a][c] += x0[a] * y0[c];
    }
    nX = sqrt(nX);
    nY = sqrt(nY);
    const double inv = 1.0 / (nX * nY);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[a][c] *= inv;
    // A = M^T M = V S^2 V^T
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) A[a][c] = M[0][a] * M[0][c] + M[1][a] * M[1][c] + M[2][a] * M[2][c];
    for (int sweep = 0; sweep < 8; ++sweep) {
      jacobi_rotate(A, V, 0, 1);
      jacobi_rotate(A, V, 0, 2);
      jacobi_rotate(A, V, 1, 2);
    }
    // order the eigenpairs by descending eigenvalue
    int o0 = 0, o1 = 1, o2 = 2;
    if (A[o0][o0] < A[o1][o1]) { const int t_ = o0; o0 = o1; o1 = t_; }
    if (A[o0][o0] < A[o2][o2]) { const int t_ = o0; o0 = o2; o2 = t_; }
    if (A[o1][o1] < A[o2][o2]) { const int t_ = o1; o1 = o2; o2 = t_; }
    const int ord[3] = {o0, o1, o2};
    double sv[3], Vs[3][3];  // Vs[:, k] = k-th right singular vector
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      sv[k] = sqrt(fmax(A[ord[k]][ord[k]], 0.0));
#pragma unroll
      for (int a = 0; a < 3; ++a) Vs[a][k] = V[a][ord[k]];
    }
    double U[3][3];  // U[:, k]
    {
      double u1[3], u2[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        u1[a] = M[a][0] * Vs[0][0] + M[a][1] * Vs[1][0] + M[a][2] * Vs[2][0];
        u2[a] = M[a][0] * Vs[0][1] + M[a][1] * Vs[1][1] + M[a][2] * Vs[2][1];
      }
      double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
      if (n1 < 1e-150) { u1[0] = 1; u1[1] = 0; u1[2] = 0; n1 = 1; }
#pragma unroll
      for (int a = 0; a < 3; ++a) u1[a] /= n1;
      const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
      for (int a = 0; a < 3; ++a) u2[a] -= d12 * u1[a];
      double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
      if (n2 < 1e-150) {  // rank-1 M (collinear joints): any unit vector orthogonal to u1
        const int m = fabs(u1[0]) <= fabs(u1[1]) ? (fabs(u1[0]) <= fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[m] = 1;
        const double de = u1[m];
#pragma unroll
        for (int a = 0; a < 3; ++a) u2[a] = e[a] - de * u1[a];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) u2[a] /= n2;
      U[0][0] = u1[0]; U[1][0] = u1[1]; U[2][0] = u1[2];
      U[0][1] = u2[0]; U[1][1] = u2[1]; U[2][1] = u2[2];
      U[0][2] = u1[1] * u2[2] - u1[2] * u2[1];
      U[1][2] = u1[2] * u2[0] - u1[0] * u2[2];
      U[2][2] = u1[0] * u2[1] - u1[1] * u2[0];
    }
    const double detV = Vs[0][0] * (Vs[1][1] * Vs[2][2] - Vs[1][2] * Vs[2][1]) -
                        Vs[0][1] * (Vs[1][0] * Vs[2][2] - Vs[1][2] * Vs[2][0]) +
                        Vs[0][2] * (Vs[1][0] * Vs[2][1] - Vs[1][1] * Vs[2][0]);
    const double detM = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
                        M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                        M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    const double sg = detV >= 0.0 ? 1.0 : -1.0;  // with det U = +1 by construction: makes det R = +1
    const double d = detM >= 0.0 ? 1.0 : -1.0;
    double R[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) R[a][c] = Vs[a][0] * U[c][0] + Vs[a][1] * U[c][1] + sg * Vs[a][2] * U[c][2];
    const double scale = (sv[0] + sv[1] + d * sv[2]) * nX / nY;
    double t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = muX[c] - scale * (muY[0] * R[0][c] + muY[1] * R[1][c] + muY[2] * R[2][c]);
    float* eo = p.err + i * kJ;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      double e2 = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double al = scale * (Y[j][0] * R[0][c] + Y[j][1] * R[1][c] + Y[j][2] * R[2][c]) + t[c];
        const double dd = al - X[j][c];
        e2 += dd * dd;
      }
      eo[j] = static_cast<float>(sqrt(e2));
    }
  }
}


}  // namespace d3dp
