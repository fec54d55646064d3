// Pipe-throughput probe for sm_100a: scalar FFMA vs packed FFMA2 vs MUFU (dev tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes tools/ubench/pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float ex2(float a) { float d; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
__device__ __forceinline__ float rcp(float a) { float d; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float a[8]; f32x2 p[8];
  for (int j = 0; j < 8; ++j) { a[j] = threadIdx.x * 1e-3f + j; p[j] = ((f32x2)__float_as_uint(a[j]) << 32) | __float_as_uint(a[j] + 1.f); }
  const f32x2 ps = ((f32x2)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) a[j] = fma1(a[j], s, a[(j + 1) & 7]);
      if (MODE == 1) p[j] = fma2(p[j], ps, p[(j + 1) & 7]);
      if (MODE == 2) a[j] = ex2(a[j]);
      if (MODE == 3) a[j] = rcp(a[j]);
      if (MODE == 4) { p[j] = fma2(p[j], ps, p[(j + 1) & 7]); a[j] = ex2(a[j]); }        // 1 FFMA2 : 1 MUFU
      if (MODE == 5) { a[j] = fma1(a[j], s, a[(j + 1) & 7]); p[j] = fma2(p[j], ps, p[(j + 1) & 7]); }
    }
  }
  float r = 0; for (int j = 0; j < 8; ++j) r += a[j] + __uint_as_float((unsigned)(p[j] >> 32)) + __uint_as_float((unsigned)p[j]);
  if (r == 12345.678f) out[0] = r;
}
template <int MODE> void run(const char* name, int per_iter) {
  float* out; cudaMalloc(&out, 4);
  const int iters = 4096, blocks = 148 * 4, threads = 512;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, iters, 0.999f);
  cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(out, iters, 0.999f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double winstr = (double)blocks * threads / 32 * iters * 8 * per_iter;
  printf("%-28s %8.3f ms  %6.2f warp-instr/ns  = %5.3f warp-instr/clk/SMSP @1.9GHz\n", name, ms, winstr / (ms * 1e6), winstr / (ms * 1e6) / (148 * 4 * 1.9));
}
int main() {
  run<0>("FFMA (3 reg)", 1); run<1>("FFMA2", 1); run<2>("MUFU.EX2", 1); run<3>("MUFU.RCP", 1); run<4>("FFMA2 + EX2", 2); run<5>("FFMA + FFMA2", 2);
  return 0;
}
