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
// The weight row is the only cold stream (DRAM latency ~0.8 us): a lane issues kFwdU 16-byte weight loads before it
// consumes the first one (a rolled loop of dependent 4-byte loads ran at one DRAM round trip per 32 input features).
constexpr int kFwdU = 6;
__global__ void __launch_bounds__(256) linear_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                             const float* __restrict__ b, float* __restrict__ pre,
                                                             float* __restrict__ y, int N, int Cin, int Cout, int act) {
  const int lane = threadIdx.x & 31, o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= Cout) return;
  const float* w = W + (long long)o * Cin;
  const float bias = b ? __ldg(b + o) : 0.0f;
  const bool vec = (Cin & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
  for (int n0 = 0; n0 < N; n0 += kRows) {
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
    if (vec) {
      const int C4 = Cin >> 2;
      const float4* w4 = reinterpret_cast<const float4*>(w);
      const float4* x4 = reinterpret_cast<const float4*>(x);
      for (int base = 0; base < C4; base += 32 * kFwdU) {
        float4 wv[kFwdU];
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) {
          const int i4 = base + u * 32 + lane;
          wv[u] = i4 < C4 ? __ldg(w4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) {
          const int i4 = base + u * 32 + lane;
          if (i4 >= C4) continue;
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            if (n0 + r < N) {
              const float4 xv = __ldg(x4 + (long long)(n0 + r) * C4 + i4);
              acc[r] = fmaf(wv[u].x, xv.x, acc[r]);
              acc[r] = fmaf(wv[u].y, xv.y, acc[r]);
              acc[r] = fmaf(wv[u].z, xv.z, acc[r]);
              acc[r] = fmaf(wv[u].w, xv.w, acc[r]);
            }
        }
      }
    } else {
      for (int base = 0; base < Cin; base += 32 * kFwdU) {
        float wv[kFwdU];
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) {
          const int i = base + u * 32 + lane;
          wv[u] = i < Cin ? __ldg(w + i) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < kFwdU; ++u) {
          const int i = base + u * 32 + lane;
          if (i >= Cin) continue;
#pragma unroll
          for (int r = 0; r < kRows; ++r)
            if (n0 + r < N) acc[r] = fmaf(wv[u], __ldg(x + (long long)(n0 + r) * Cin + i), acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], s);
    }
    if (lane == 0) {
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
  linear_act_fwd_kernel<<<(Cout + 7) / 8, 256, 0, ST>>>(x, W, b, pre, y, N, Cin, Cout, act);
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
