// [N, <=1536]-sized fp32 linear layers of the hot path: the SEBlock gate MLP (new_scripy.py:148-152: Linear - GELU -
// Linear - Sigmoid on the pooled vector) and EmbedFC (new_scripy.py:255-268: Linear - GELU - Linear on t / the masked
// one-hot class).  GEMV-class work (N = 4 training rows, <= 30 sampling rows): every kernel is weight-stationary -- a
// warp / thread streams its slice of the weight matrix once with coalesced loads while the few activation rows sit in
// L1 / shared memory -- and deterministic (no atomics: the input gradient leaves as per-slice partial rows that the
// next kernel in the chain folds while it applies the activation derivative).
#include "common.cuh"
#include "dm_b200.h"

namespace {

constexpr int kRows = 8;        // activation rows per pass
constexpr int kSlice = 32;      // output features per backward block (= partial rows of the input gradient)
constexpr int kBwdThreads = 128;

__device__ __forceinline__ float lin_act(float v, int act) {
  if (act == 1) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  if (act == 2) return fmaxf(v, 0.0f);
  if (act == 3) return 1.0f / (1.0f + expf(-v));
  return v;
}
// derivative from aux = pre-activation (gelu, relu) or output (sigmoid)
__device__ __forceinline__ float lin_act_grad(float aux, int act) {
  if (act == 1) {
    const float cdf = 0.5f * (1.0f + erff(aux * 0.70710678118654752f));
    return cdf + aux * 0.3989422804014327f * expf(-0.5f * aux * aux);
  }
  if (act == 2) return aux > 0.0f ? 1.0f : 0.0f;
  if (act == 3) return aux * (1.0f - aux);
  return 1.0f;
}

// y[n,o] = act(b[o] + sum_i W[o,i] x[n,i]).  One warp per output feature, lanes stride the input features.
// Two streams, both requested in one batch: the warp's weight row (cold, DRAM latency ~0.8 us) as kFwdB 16-byte loads per
// lane issued before anything else, and the kRows activation rows staged once per block in shared memory while those
// loads fly (read per use from L1/L2 they came back in groups of two behind each other, 15 us for a 9.4 MB matrix; a
// rolled loop of dependent 4-byte weight loads before that ran at one DRAM round trip per 32 input features, 33 us).
constexpr int kFwdB = 12;                  // weight loads in flight per lane
constexpr int kChunk = 32 * kFwdB * 4;     // input features per staged chunk (1536: 48 KB of rows)

template <bool VEC>
__device__ __forceinline__ void lin_fwd_rows(const float* __restrict__ x, const float* __restrict__ w, float* xs,
                                             float (&acc)[kRows], int n0, int N, int Cin, bool live, int lane) {
  constexpr int VW = VEC ? 4 : 1, kSpan = 32 * kFwdB * VW;
  for (int c0 = 0; c0 < Cin; c0 += kChunk) {
    const int cn = min(kChunk, Cin - c0);
    float wv[kFwdB][VW];
    auto load_batch = [&](int b0) {
#pragma unroll
      for (int u = 0; u < kFwdB; ++u) {
        const int j = b0 + (u * 32 + lane) * VW;
        const bool ok = live && j < cn;
        if (VEC) {
          const float4 t = ok ? __ldg(reinterpret_cast<const float4*>(w + c0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          wv[u][0] = t.x; wv[u][VW > 1 ? 1 : 0] = t.y; wv[u][VW > 2 ? 2 : 0] = t.z; wv[u][VW > 3 ? 3 : 0] = t.w;
        } else {
          wv[u][0] = ok ? __ldg(w + c0 + j) : 0.0f;
        }
      }
    };
    load_batch(0);
    __syncthreads();                                   // the previous chunk's readers are done
#pragma unroll
    for (int r = 0; r < kRows; ++r) {                  // rows past N are staged as zeros: no row test in the FMA loop
      const bool row = n0 + r < N;
      const float* xr = x + (long long)(n0 + r) * Cin + c0;
      if (VEC) {
        for (int j4 = threadIdx.x; j4 < (cn >> 2); j4 += blockDim.x)
          *reinterpret_cast<float4*>(xs + r * kChunk + 4 * j4) =
              row ? __ldg(reinterpret_cast<const float4*>(xr) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        for (int j = threadIdx.x; j < cn; j += blockDim.x) xs[r * kChunk + j] = row ? __ldg(xr + j) : 0.0f;
      }
    }
    __syncthreads();
    for (int b0 = 0; b0 < cn; b0 += kSpan) {
      if (b0) load_batch(b0);
#pragma unroll
      for (int u = 0; u < kFwdB; ++u) {
        const int j = b0 + (u * 32 + lane) * VW;
        if (j >= cn) continue;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
          if (VEC) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + r * kChunk + j);
            acc[r] = fmaf(wv[u][0], xv.x, acc[r]);
            acc[r] = fmaf(wv[u][VW > 1 ? 1 : 0], xv.y, acc[r]);
            acc[r] = fmaf(wv[u][VW > 2 ? 2 : 0], xv.z, acc[r]);
            acc[r] = fmaf(wv[u][VW > 3 ? 3 : 0], xv.w, acc[r]);
          } else {
            acc[r] = fmaf(wv[u][0], xs[r * kChunk + j], acc[r]);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) linear_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                             const float* __restrict__ b, float* __restrict__ pre,
                                                             float* __restrict__ y, int N, int Cin, int Cout, int act) {
  extern __shared__ __align__(16) float xs[];          // [kRows][kChunk]
  const int lane = threadIdx.x & 31, o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const bool live = o < Cout;                          // dead warps still help staging and hit the barriers
  const float* w = W + (long long)(live ? o : 0) * Cin;
  const float bias = (b && live) ? __ldg(b + o) : 0.0f;
  const bool vec = (Cin & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
  for (int n0 = 0; n0 < N; n0 += kRows) {
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
    if (vec) lin_fwd_rows<true>(x, w, xs, acc, n0, N, Cin, live, lane);
    else lin_fwd_rows<false>(x, w, xs, acc, n0, N, Cin, live, lane);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
    }
    if (lane == 0 && live) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (n0 + r < N) {
          const float v = acc[r] + bias;
          if (pre) pre[(long long)(n0 + r) * Cout + o] = v;
          y[(long long)(n0 + r) * Cout + o] = lin_act(v, act);
        }
    }
  }
}

// g[n,o] = (sum_p dy[p][n][o]) * act'(aux[n,o]) for the block's 32 output features; then per input feature i (one
// thread each): dW[o,i] += sum_n g[n,o] x[n,i],  dx_parts[slice][n][i] = sum_{o in slice} g[n,o] W[o,i];  the blocks of
// the first input tile also add db[o] += sum_n g[n,o].
__global__ void __launch_bounds__(kBwdThreads) linear_act_bwd_kernel(const float* __restrict__ dy, int nparts,
                                                                     const float* __restrict__ aux, int act,
                                                                     const float* __restrict__ x, const float* __restrict__ W,
                                                                     float* __restrict__ dW, float* __restrict__ db,
                                                                     float* __restrict__ dx_parts, int N, int Cin, int Cout) {
  __shared__ float g[kRows][kSlice];
  const int i = blockIdx.x * kBwdThreads + threadIdx.x, o0 = blockIdx.y * kSlice;
  const int no = min(kSlice, Cout - o0);
  const bool live = i < Cin;
  float dwacc[kSlice];
#pragma unroll
  for (int o = 0; o < kSlice; ++o) dwacc[o] = 0.0f;
  float dbacc = 0.0f;
  const long long part_stride = (long long)N * Cout;
  // The accumulate-into-.grad reads are issued first, before anything depends on them, so their DRAM round trip hides
  // behind the two phases below.  (Written as `dW[..] += ..` at the end, the compiler has to keep every load behind the
  // previous store -- it cannot prove the rows distinct -- which cost 32 serial DRAM round trips per thread.)
  float old[kSlice];
#pragma unroll
  for (int o = 0; o < kSlice; ++o) old[o] = (live && dW && o < no) ? __ldcg(dW + (long long)(o0 + o) * Cin + i) : 0.0f;
  for (int n0 = 0; n0 < N; n0 += kRows) {
    __syncthreads();
    for (int e = threadIdx.x; e < kRows * kSlice; e += kBwdThreads) {
      const int r = e / kSlice, o = e % kSlice;
      float v = 0.0f;
      if (n0 + r < N && o < no) {
        const long long at = (long long)(n0 + r) * Cout + o0 + o;
        for (int p0 = 0; p0 < nparts; p0 += 8) {       // eight partial rows in flight, folded in row order
          float t[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) t[k] = p0 + k < nparts ? __ldg(dy + (p0 + k) * part_stride + at) : 0.0f;
#pragma unroll
          for (int k = 0; k < 8; ++k) v += t[k];
        }
        v *= lin_act_grad(act ? __ldg(aux + at) : 0.0f, act);
      }
      g[r][o] = v;
    }
    __syncthreads();
    if (db && blockIdx.x == 0 && threadIdx.x < kSlice) {
#pragma unroll
      for (int r = 0; r < kRows; ++r) dbacc += g[r][threadIdx.x];
    }
    if (!live) continue;
    float xv[kRows], dxacc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      xv[r] = (n0 + r < N) ? __ldg(x + (long long)(n0 + r) * Cin + i) : 0.0f;
      dxacc[r] = 0.0f;
    }
#pragma unroll
    for (int o = 0; o < kSlice; ++o) {
      const float wv = (dx_parts && o < no) ? __ldg(W + (long long)(o0 + o) * Cin + i) : 0.0f;
#pragma unroll
      for (int r = 0; r < kRows; ++r) {
        const float gv = g[r][o];
        dwacc[o] = fmaf(gv, xv[r], dwacc[o]);
        dxacc[r] = fmaf(gv, wv, dxacc[r]);
      }
    }
    if (dx_parts) {
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (n0 + r < N) dx_parts[((long long)blockIdx.y * N + n0 + r) * Cin + i] = dxacc[r];
    }
  }
  if (db && blockIdx.x == 0 && threadIdx.x < no) db[o0 + threadIdx.x] += dbacc;
  if (live && dW) {
#pragma unroll
    for (int o = 0; o < kSlice; ++o)
      if (o < no) dW[(long long)(o0 + o) * Cin + i] = old[o] + dwacc[o];
  }
}

__global__ void __launch_bounds__(256) sum_parts_kernel(const float* __restrict__ parts, int nparts, float* __restrict__ out,
                                                        long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = 0.0f;
  for (int p0 = 0; p0 < nparts; p0 += 8) {
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = p0 + k < nparts ? __ldg(parts + (p0 + k) * n + i) : 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += t[k];
  }
  out[i] = v;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int dm_linear_act_fwd(const float* x, const float* W, const float* b, float* pre, float* y, int N, int Cin,
                                 int Cout, int act, void* stream) {
  if (act < 0 || act > 3) { dm_set_error("dm_linear_act_fwd: act must be 0 (none), 1 (gelu), 2 (relu) or 3 (sigmoid)"); return DM_ERR_ARG; }
  if (N <= 0 || Cout <= 0) return DM_OK;
  if (Cin <= 0) { dm_set_error("dm_linear_act_fwd: Cin must be positive"); return DM_ERR_ARG; }
  linear_act_fwd_kernel<<<(Cout + 7) / 8, 256, (size_t)kRows * kChunk * sizeof(float), ST>>>(x, W, b, pre, y, N, Cin, Cout, act);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

extern "C" int dm_linear_bwd_parts(int Cout) { return (Cout + kSlice - 1) / kSlice; }

extern "C" int dm_linear_act_bwd(const float* dy, int nparts, const float* aux, int act, const float* x, const float* W,
                                 float* dW, float* db, float* dx_parts, int N, int Cin, int Cout, void* stream) {
  if (act < 0 || act > 3) { dm_set_error("dm_linear_act_bwd: act must be 0 (none), 1 (gelu), 2 (relu) or 3 (sigmoid)"); return DM_ERR_ARG; }
  if (act && !aux) { dm_set_error("dm_linear_act_bwd: aux (pre-activation, or the output for sigmoid) is required"); return DM_ERR_ARG; }
  if (nparts < 1) { dm_set_error("dm_linear_act_bwd: nparts must be >= 1"); return DM_ERR_ARG; }
  if (N <= 0 || Cout <= 0 || Cin <= 0) return DM_OK;
  dim3 grid((Cin + kBwdThreads - 1) / kBwdThreads, (Cout + kSlice - 1) / kSlice);
  linear_act_bwd_kernel<<<grid, kBwdThreads, 0, ST>>>(dy, nparts, aux, act, x, W, dW, db, dx_parts, N, Cin, Cout);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

extern "C" int dm_sum_parts(const float* parts, int nparts, float* out, long long n, void* stream) {
  if (n <= 0) return DM_OK;
  if (nparts < 1) { dm_set_error("dm_sum_parts: nparts must be >= 1"); return DM_ERR_ARG; }
  sum_parts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST>>>(parts, nparts, out, n);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

// Masked one-hot class rows, the input of the context EmbedFCs (new_scripy.py:337-340: one_hot(c) * ctx_mask, no flip;
// MNIST_script.py:165-171: one_hot(c) * -(1 - context_mask)).  mask_i64 != 0: the mask holds int64 (DDPM.sample builds it
// with zeros_like(c_i)), else fp32.  One launch instead of one_hot + type + repeat + mul (+ the flip).
namespace {
__global__ void ctx_onehot_kernel(const long long* __restrict__ c, const void* __restrict__ mask, int mask_i64, float* out,
                                  int N, int ncls, int flip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * ncls) return;
  const int n = i / ncls, k = i - n * ncls;
  float m = mask_i64 ? (float)reinterpret_cast<const long long*>(mask)[n] : reinterpret_cast<const float*>(mask)[n];
  if (flip) m = -1.0f * (1.0f - m);
  out[i] = (c[n] == k ? 1.0f : 0.0f) * m;
}
}  // namespace

extern "C" int dm_ctx_onehot(const long long* c, const void* mask, int mask_i64, float* out, int N, int ncls, int flip,
                             void* stream) {
  if (N <= 0 || ncls <= 0) { dm_set_error("dm_ctx_onehot: bad sizes"); return DM_ERR_ARG; }
  const int total = N * ncls;
  ctx_onehot_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, mask, mask_i64, out, N, ncls, flip);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
