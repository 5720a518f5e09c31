// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of warps and the access shape.
// One CTA per SM (grid = #SMs, only CTA 0 reports), 512 TMEM columns allocated, every warp reads its own lane quadrant
// (warp % 4) REPS times back to back and waits once at the end.  Reports bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../d3dp_b200/csrc/ptx.cuh"
using namespace d3dp;

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 256 bits per repetition: .16x256b.x8 = 16 rows x 64 columns?  (x1 = 4 regs/thread)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// mode 0: ld 32x32b.x32 ; 1: ld 32x32b.x16 ; 2: ld 32x32b.x64 ; 3: ld 16x256b.x8 ; 4: st 32x32b.x32 ;
// 5: ld x32 interleaved with 32 MUFU.EX2 + 32 FFMA per load (overlap test) ; 6: the same math without loads
template <int MODE>
__global__ void __launch_bounds__(512, 1) bench(int reps, long long* out, float* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = threadIdx.x + i;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    const uint32_t col = (r * 64) & 255;  // stay inside the allocation
    if constexpr (MODE == 0) tmem_ld32(base + col, *reinterpret_cast<uint32_t(*)[32]>(v));
    if constexpr (MODE == 1) tmem_ld16(base + col, v);
    if constexpr (MODE == 2) tmem_ld64(base + col, v);
    if constexpr (MODE == 3) tmem_ld_16x256b_x8(base + col, v);
    if constexpr (MODE == 4) tmem_st32(base + col, *reinterpret_cast<uint32_t(*)[32]>(v));
    if constexpr (MODE == 5 || MODE == 6) {
      uint32_t w[32];
      if (MODE == 5) tmem_ld32(base + col, w);
      // math on the PREVIOUS chunk (v) while the load is in flight
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += ex2_approx(fmaf(__uint_as_float(v[i]), 0.18f, -1.0f));
      if (MODE == 5) {
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = w[i] & 0x3fffffffu;
      }
    }
  }
  if (MODE == 4) tmem_st_wait(); else tmem_ld_wait();
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) acc += __uint_as_float(v[i]);
  if (acc == 12345.678f) sink[0] = acc;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_ptr);
}

template <int MODE>
void run(const char* name, int bytes_per_warp_op, int sms) {
  long long* d;
  float* s;
  cudaMalloc(&d, 8);
  cudaMalloc(&s, 4);
  const int reps = 2000;
  for (int warps : {1, 2, 4, 8, 16}) {
    bench<MODE><<<sms, warps * 32>>>(reps, d, s);
    bench<MODE><<<sms, warps * 32>>>(reps, d, s);
    long long clk = 0;
    cudaError_t e = cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s warps=%d ERROR %s\n", name, warps, cudaGetErrorString(e)); return; }
    const double bpc = bytes_per_warp_op ? double(bytes_per_warp_op) * reps * warps / clk : 0.0;
    printf("%-34s warps=%2d  clk=%9lld  clk/op/warp=%7.1f  bytes/clk/SM=%7.1f\n", name, warps, clk, double(clk) / reps, bpc);
  }
  cudaFree(d);
  cudaFree(s);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s  SMs=%d  clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("ld 32x32b.x32 (4 KB/warp)", 4096, p.multiProcessorCount);
  run<1>("ld 32x32b.x16 (2 KB/warp)", 2048, p.multiProcessorCount);
  run<2>("ld 32x32b.x64 (8 KB/warp)", 8192, p.multiProcessorCount);
  run<3>("ld 16x256b.x8 (4 KB/warp)", 4096, p.multiProcessorCount);
  run<4>("st 32x32b.x32 (4 KB/warp)", 4096, p.multiProcessorCount);
  run<5>("ld x32 + 32 ex2/fma per thread", 4096, p.multiProcessorCount);
  run<6>("32 ex2/fma per thread, no ld", 0, p.multiProcessorCount);
  return 0;
}
