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
void dm_note_kernel(const char* name, int param);   // which kernel variant a dispatcher chose (dm_kernel_count / dm_last_kernel)
long long dm_debug_value(int key);      // dev switches set through dm_debug_set (conv_gemm.cu)

#define DM_NUM_SMS 148

namespace dm {

typedef __nv_bfloat16 bf16;

struct __align__(16) bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// one 16-byte load / store (a struct copy of four bf16x2 compiles to four 4-byte accesses)
__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  f[0] = bf_lo(r.x); f[1] = bf_hi(r.x); f[2] = bf_lo(r.y); f[3] = bf_hi(r.y);
  f[4] = bf_lo(r.z); f[5] = bf_hi(r.z); f[6] = bf_lo(r.w); f[7] = bf_hi(r.w);
}
// Batched streaming loads: `volatile` keeps a group of these in program order, so the N loads of an unrolled
// "load N pixels, then do the math" loop really are in flight together (left alone, ptxas may interleave each
// load with the previous one's consumers and serialise the round trips).
__device__ __forceinline__ uint4 ldg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Per-thread asynchronous 16-byte copies global -> shared (LDGSTS): the memory-level parallelism of a streaming
// kernel then no longer depends on how ptxas schedules register loads against their consumers.  `pred` false
// zero-fills the slot without touching global memory.  A thread only ever reads back its own slots, so
// cp.async.wait_group is the only synchronisation needed.
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* g, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds16(uint32_t smem_addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_addr) : "memory");
  return r;
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  f[0] = bf_lo(r.x); f[1] = bf_hi(r.x); f[2] = bf_lo(r.y); f[3] = bf_hi(r.y);
  f[4] = bf_lo(r.z); f[5] = bf_hi(r.z); f[6] = bf_lo(r.w); f[7] = bf_hi(r.w);
}
__device__ __forceinline__ uint32_t pack2(float a, float b);
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 r;
  r.x = pack2(f[0], f[1]); r.y = pack2(f[2], f[3]); r.z = pack2(f[4], f[5]); r.w = pack2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = r;
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
// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): one issue slot does two lanes' worth of work.  The
// streaming norm kernels are issue-bound on the GELU evaluation, so everything that is not a MUFU or an integer
// op is done on register pairs; constants broadcast as immediates.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f32x2 bc2(float c) { return pk2(c, c); }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// bf16x2 word <-> packed pair (element 0 in the low half)
__device__ __forceinline__ f32x2 bf2_to_f2(uint32_t w) { return pk2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f2_to_bf2(f32x2 v) { float a, b; upk2(v, a, b); return pack2(a, b); }

// Abramowitz-Stegun 7.1.26 on pairs, in the variable u (x = u/sqrt2 folded into the constants), coefficients
// halved: poly(t) * t * exp(-u^2/2) = 0.5*erfc(|u|/sqrt2) = Phi(-|u|).
constexpr float kAsP = 0.3275911f * 0.70710678118654752f;
struct AsPair { f32x2 poly, T, D, E; };       // poly(t) (without the final *t), t, d = 1 + p|u|, exp(-u^2/2)
__device__ __forceinline__ AsPair as_pair(f32x2 U, float u0, float u1) {
  AsPair r;
  const float d0 = fmaf(fabsf(u0), kAsP, 1.0f), d1 = fmaf(fabsf(u1), kAsP, 1.0f);
  r.D = pk2(d0, d1);
  r.T = pk2(rcp_approx(d0), rcp_approx(d1));
  f32x2 poly = fma2(r.T, bc2(0.5307027145f), bc2(-0.7265760135f));
  poly = fma2(poly, r.T, bc2(0.7107068705f));
  poly = fma2(poly, r.T, bc2(-0.142248368f));
  r.poly = fma2(poly, r.T, bc2(0.127414796f));
  float m0, m1;
  upk2(mul2(mul2(U, U), bc2(-0.72134752044448170f)), m0, m1);
  r.E = pk2(ex2_approx(m0), ex2_approx(m1));
  return r;
}
// gelu(u) = u*Phi(u) = max(u,0) - |u| * Phi(-|u|)
__device__ __forceinline__ f32x2 gelu2(f32x2 U) {
  float u0, u1; upk2(U, u0, u1);
  const AsPair a = as_pair(U, u0, u1);
  float q0, q1; upk2(mul2(mul2(a.poly, a.T), a.E), q0, q1);
  return pk2(fmaf(-fabsf(u0), q0, fmaxf(u0, 0.0f)), fmaf(-fabsf(u1), q1, fmaxf(u1, 0.0f)));
}
// gelu'(u) = Phi(u) + u*phi(u).  With r = exp(-u^2/2) * (poly(t)*t - c|u|), c = 1/sqrt(2 pi):  gelu' = r for u < 0
// and 1 - r for u >= 0, i.e. 0.5 + copysign(0.5 - r, u) (0.5 - r >= 0 everywhere).  c|u| = (c/p)*(d - 1).
__device__ __forceinline__ f32x2 gelu_grad2(f32x2 U) {
  float u0, u1; upk2(U, u0, u1);
  const AsPair a = as_pair(U, u0, u1);
  constexpr float kc = 0.3989422804014327f / kAsP;
  const f32x2 V = fma2(a.D, bc2(-kc), fma2(a.poly, a.T, bc2(kc)));
  float h0, h1; upk2(fma2(mul2(V, a.E), bc2(-1.0f), bc2(0.5f)), h0, h1);
  h0 = __uint_as_float((__float_as_uint(h0) & 0x7fffffffu) | (__float_as_uint(u0) & 0x80000000u));
  h1 = __uint_as_float((__float_as_uint(h1) & 0x7fffffffu) | (__float_as_uint(u1) & 0x80000000u));
  return add2(pk2(h0, h1), bc2(0.5f));
}
template <int ACT> __device__ __forceinline__ f32x2 act2(f32x2 U) {        // ACT: 0 none, 1 GELU, 2 ReLU
  if (ACT == 1) return gelu2(U);
  if (ACT == 2) { float a, b; upk2(U, a, b); return pk2(fmaxf(a, 0.0f), fmaxf(b, 0.0f)); }
  return U;
}
template <int ACT> __device__ __forceinline__ f32x2 act_grad2(f32x2 U) {
  if (ACT == 1) return gelu_grad2(U);
  if (ACT == 2) { float a, b; upk2(U, a, b); return pk2(a > 0.0f ? 1.0f : 0.0f, b > 0.0f ? 1.0f : 0.0f); }
  return bc2(1.0f);
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
