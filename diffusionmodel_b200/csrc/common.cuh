// Shared device helpers for the sm_100a diffusion hot-path kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define DM_OK 0
#define DM_ERR_ARG (-1)
#define DM_ERR_CUDA (-2)
#define DM_ERR_TMA (-3)

#define DM_CHECK_LAUNCH()                                      \
  do {                                                         \
    dm_count_launch();                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) { dm_set_error(cudaGetErrorString(e__)); return DM_ERR_CUDA; } \
  } while (0)

void dm_set_error(const char* msg);
void dm_count_launch();

#define DM_NUM_SMS 148

namespace dm {

typedef __nv_bfloat16 bf16;

struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  bf16x8 r = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(r.v[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  bf16x8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<bf16x8*>(p) = r;
}
__device__ __forceinline__ void load4(const bf16* p, float (&f)[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void store4(bf16* p, const float (&f)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// GELU (nn.GELU() default = exact erf) and its derivative.  erf by Abramowitz-Stegun 7.1.26 (|abs err| <
// 1.5e-7 in exact arithmetic, < 5e-7 as evaluated here in fp32: three orders below the bf16 rounding of
// the stored activation) -- one MUFU.RCP + one MUFU.EX2 instead of the ~30-instruction erff(), and the
// same exponential exp(-x^2/2) serves the Gaussian density in the derivative.
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));      // one MUFU.RCP, no slow-path call
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
struct GeluParts { float cdf, pdf_e; };   // Phi(x), exp(-x^2/2)
__device__ __forceinline__ GeluParts gelu_parts(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = ex2_approx(x * x * -0.72134752044448170f);   // exp(-x^2/2) = 2^(-x^2 * log2(e)/2)
  const float half_erfc = 0.5f * poly * t * e;             // 0.5 * erfc(|x|/sqrt2)
  GeluParts r;
  r.cdf = x >= 0.0f ? 1.0f - half_erfc : half_erfc;
  r.pdf_e = e;
  return r;
}
__device__ __forceinline__ float gelu_f(float x) { return x * gelu_parts(x).cdf; }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const GeluParts g = gelu_parts(x);
  return fmaf(x * 0.3989422804014327f, g.pdf_e, g.cdf);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// act: 0 none, 1 GELU, 2 ReLU
__device__ __forceinline__ float act_f(float x, int act) {
  return act == 1 ? gelu_f(x) : (act == 2 ? fmaxf(x, 0.0f) : x);
}
__device__ __forceinline__ float act_grad_f(float x, int act) {
  return act == 1 ? gelu_grad_f(x) : (act == 2 ? (x > 0.0f ? 1.0f : 0.0f) : 1.0f);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of one float; result valid in every thread.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  float r = (lane < nw) ? smem32[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace dm
