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
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// exact-erf GELU (nn.GELU() default) and its derivative
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
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
