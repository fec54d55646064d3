// Bandwidth-bound kernels of the diffusion hot path: normalisation + activation, squeeze-excite and
// coordinate-attention pooling/gating, FiLM, bilinear upsample + concat, pooling, LocalEnhancer mask
// weighting, layout conversion.  All activations are bf16 NHWC with a channel pitch; every thread
// moves 16-byte vectors (8 channels) and all reductions are fp32 (warp shuffles + shared memory).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "dm_b200.h"

using dm::bf16;
using dm::load8;
using dm::store8;

namespace {

constexpr int kEwThreads = 256;

inline int ew_grid(long long total) {
  long long b = (total + kEwThreads - 1) / kEwThreads;
  long long cap = (long long)DM_NUM_SMS * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Generic (pixel, 8-channel vector) elementwise driver.  F::operator()(p, c0) handles channels [c0, c0+8).
template <class F>
__global__ void __launch_bounds__(kEwThreads) ew_kernel(unsigned total, unsigned Cv, F f) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned p = i / Cv, cv = i - p * Cv;
    f(p, (int)(cv * 8));
  }
}
template <class F>
int ew_launch(long long P, int C, F f, cudaStream_t st) {
  long long Cv = (C + 7) / 8, total = P * Cv;
  if (total <= 0) return DM_OK;
  if (total >= (1ll << 32)) { dm_set_error("elementwise: tensor too large"); return DM_ERR_ARG; }
  ew_kernel<F><<<ew_grid(total), kEwThreads, 0, st>>>((unsigned)total, (unsigned)Cv, f);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

__device__ __forceinline__ void ldp8(const float* p, int c0, int C, float (&o)[8]) {
  if (c0 + 8 <= C && ((reinterpret_cast<uintptr_t>(p + c0) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + c0)), b = __ldg(reinterpret_cast<const float4*>(p + c0) + 1);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (c0 + j < C) ? __ldg(p + c0 + j) : 0.0f;
}

// Thread mapping of the channel-parameterised streaming kernels (BatchNorm apply / backward): a thread
// keeps ONE 8-channel vector for its whole life, so the per-channel coefficients are loaded once into
// registers and the inner loop is pure 16-byte loads -> math -> 16-byte stores.  A block is
// VPB channel-vectors x R pixel rows; consecutive threads touch consecutive 16-byte chunks of a pixel.
struct ChanMap { int Cv, VPB, R, cvt, threads; };
inline ChanMap chan_map(int C, int vec = 8) {
  ChanMap m;
  m.Cv = (C + vec - 1) / vec;
  m.VPB = m.Cv < 256 ? m.Cv : 256;
  m.R = 256 / m.VPB; if (m.R < 1) m.R = 1;
  m.cvt = (m.Cv + m.VPB - 1) / m.VPB;
  m.threads = m.VPB * m.R;
  return m;
}
inline int chan_grid_x(long long P, const ChanMap& m, int px_per_thread) {
  long long want = (P + (long long)m.R * px_per_thread - 1) / ((long long)m.R * px_per_thread);
  long long cap = (long long)DM_NUM_SMS * 8 / m.cvt; if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// ---------------------------------------------------------------------------------- reductions
// out1[g][c] (+ out2[g][c]) += scale * sum over the `count` pixels of group g.
struct RedArgs {
  const bf16* a; const bf16* b;
  long long a_hi, a_lo, a_ps, b_hi, b_lo, b_ps;   // group base = (g/gdiv)*hi + (g%gdiv)*lo ; pixel stride ps
  int gdiv, count, C, mode, act, G, stat_div;      // stat index for mode 4: (g / stat_div) * G + c / (C/G)
  const float *mean, *invstd, *gamma, *beta;
  float scale;
  float *out1, *out2;
  float* part;          // workspace for the per-split partial sums (splits > 1)
};
// MODE 0: s1 = sum a            MODE 2: s1 = sum a*b, s2 = sum a
// Four pixels' 16-byte loads in flight per thread and operand.
template <int MODE>
__global__ void __launch_bounds__(256) nc_reduce_kernel(RedArgs A) {
  __shared__ float sm[256 * 16];
  const int Cv = (A.C + 7) / 8;
  const int VPB = Cv < 256 ? Cv : 256;
  const int R = 256 / VPB;
  const int t = threadIdx.x, cvl = t % VPB, r = t / VPB;
  const int cv = blockIdx.y * VPB + cvl;
  const int g = blockIdx.z;
  const int c0 = cv * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const bool live = (r < R) && (cv < Cv);
  if (live) {
    const bf16* pa = A.a + (long long)(g / A.gdiv) * A.a_hi + (long long)(g % A.gdiv) * A.a_lo + c0;
    const bf16* pb = MODE == 2 ? A.b + (long long)(g / A.gdiv) * A.b_hi + (long long)(g % A.gdiv) * A.b_lo + c0 : nullptr;
    const int chunk = (A.count + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * chunk;
    const int i1 = min(A.count, i0 + chunk);
    constexpr int NB = MODE == 2 ? 4 : 1;
    uint4 wa[4], wb[NB];
    auto fetch = [&](int i, uint4 (&xa)[4], uint4 (&xb)[NB]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ii = i + u * R;
        const bool ok = ii < i1;
        xa[u] = ok ? dm::ldg16(pa + (long long)ii * A.a_ps) : make_uint4(0u, 0u, 0u, 0u);
        if (MODE == 2) xb[MODE == 2 ? u : 0] = ok ? dm::ldg16(pb + (long long)ii * A.b_ps) : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    int i = i0 + r;
    fetch(i, wa, wb);
    while (i < i1) {                                  // software pipeline, see bn_stats_kernel
      uint4 na[4], nb[NB];
      fetch(i + 4 * R, na, nb);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float va[8], vb[8];
        dm::unpack8(wa[u], va);
        if (MODE == 2) dm::unpack8(wb[MODE == 2 ? u : 0], vb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) s1[j] += va[j];
          else { s1[j] = fmaf(va[j], vb[j], s1[j]); s2[j] += va[j]; }
        }
        wa[u] = na[u];
        if (MODE == 2) wb[MODE == 2 ? u : 0] = nb[MODE == 2 ? u : 0];
      }
      i += 4 * R;
    }
  }
  // reduce over the R pixel rows of the block
#pragma unroll
  for (int j = 0; j < 8; ++j) { sm[t * 16 + j] = s1[j]; sm[t * 16 + 8 + j] = s2[j]; }
  __syncthreads();
  if (r == 0 && cv < Cv) {
    for (int rr = 1; rr < R; ++rr) {
      const float* o = sm + (rr * VPB + cvl) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += o[j]; s2[j] += o[8 + j]; }
    }
    // no atomics: a single split writes the result, several splits write partial rows that
    // reduce_finalize_kernel adds in a fixed order (bit-reproducible, and no same-address contention)
    if (gridDim.x == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < A.C) {
          A.out1[(long long)g * A.C + c0 + j] = s1[j] * A.scale;
          if (A.out2) A.out2[(long long)g * A.C + c0 + j] = s2[j] * A.scale;
        }
    } else {
      float* row = A.part + ((long long)blockIdx.x * gridDim.z + g) * 2 * A.C;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < A.C) { row[c0 + j] = s1[j]; row[A.C + c0 + j] = s2[j]; }
    }
  }
}
// Second stage of the deterministic reductions: out[i] = scale * sum_sp part[sp][i] over `splits` partial rows.
// 32 outputs x 32 row slices per block; each thread issues its rows' loads four at a time (a plain dependent
// loop pays one L2 round trip per partial row -- 130 us for the 296 rows of a column sum).  Fixed summation
// order => bit-reproducible.
__global__ void __launch_bounds__(1024) reduce_finalize_kernel(const float* __restrict__ part, int splits, int groups, int C,
                                                                float scale, float* __restrict__ out1, float* __restrict__ out2,
                                                                int accumulate) {
  __shared__ float sa[64][17], sb[64][17];
  const int cl = threadIdx.x & 15, sl = threadIdx.x >> 4;          // 16 outputs x 64 row slices
  const long long total = (long long)groups * C;
  const long long i = (long long)blockIdx.x * 16 + cl;
  float a = 0.f, b = 0.f, o1 = 0.f, o2 = 0.f;
  if (accumulate && sl == 0 && i < total) { o1 = __ldcg(out1 + i); if (out2) o2 = __ldcg(out2 + i); }   // old values first
  if (i < total) {
    const long long g = i / C; const int c = (int)(i - g * C);
    const float* base = part + g * 2 * C + c;
    const long long stride = (long long)groups * 2 * C;
    for (int sp = sl; sp < splits; sp += 512) {
      float x[8], y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int s = sp + 64 * u;
        const bool ok = s < splits;
        const float* row = base + (long long)(ok ? s : sp) * stride;
        x[u] = ok ? __ldg(row) : 0.f;
        y[u] = (ok && out2) ? __ldg(row + C) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { a += x[u]; b += y[u]; }
    }
  }
  sa[sl][cl] = a; sb[sl][cl] = b;
  __syncthreads();
  if (sl == 0 && i < total) {
    for (int k = 1; k < 64; ++k) { a += sa[k][cl]; b += sb[k][cl]; }
    out1[i] = o1 + a * scale;
    if (out2) out2[i] = o2 + b * scale;
  }
}
static inline int finalize_grid(long long total) { return (int)((total + 15) / 16); }
// few partial rows (<= 8) but possibly many outputs (CoordAttn row / column means): one thread per output
__global__ void __launch_bounds__(256) reduce_finalize_small_kernel(const float* __restrict__ part, int splits, long long total,
                                                                     int C, float scale, float* __restrict__ out1,
                                                                     float* __restrict__ out2, int accumulate) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const long long g = i / C; const int c = (int)(i - g * C);
  const float* base = part + g * 2 * C + c;
  const long long stride = (total / C) * 2 * C;
  float x[8], y[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const bool ok = u < splits;
    x[u] = ok ? __ldg(base + u * stride) : 0.f;
    y[u] = (ok && out2) ? __ldg(base + u * stride + C) : 0.f;
  }
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u) { a += x[u]; b += y[u]; }
  if (accumulate) { out1[i] += a * scale; if (out2) out2[i] += b * scale; }
  else { out1[i] = a * scale; if (out2) out2[i] = b * scale; }
}

float* g_ws = nullptr;          // caller-provided scratch for partial sums (dm_set_workspace)
long long g_ws_floats = 0;

int launch_reduce(RedArgs& A, int groups, cudaStream_t st) {
  const int Cv = (A.C + 7) / 8;
  const int VPB = Cv < 256 ? Cv : 256;
  const int R = 256 / VPB;
  const int cvt = (Cv + VPB - 1) / VPB;
  long long want = (long long)DM_NUM_SMS * 6 / ((long long)groups * cvt);      // ~1400 resident threads per SM
  int maxs = A.count / (R * 8); if (maxs < 1) maxs = 1;
  if (maxs > 512) maxs = 512;
  int splits = (int)(want < 1 ? 1 : (want > maxs ? maxs : want));
  if (groups > 65535) { dm_set_error("reduce: too many groups"); return DM_ERR_ARG; }
  if (splits > 1 && (long long)splits * groups * 2 * A.C > g_ws_floats) {
    if (g_ws_floats == 0) { dm_set_error("reduction kernels need scratch: call dm_set_workspace() first"); return DM_ERR_ARG; }
    splits = (int)(g_ws_floats / ((long long)groups * 2 * A.C));
    if (splits < 1) splits = 1;
  }
  A.part = g_ws;
  dim3 grid(splits, cvt, groups);
  if (A.mode == 2) nc_reduce_kernel<2><<<grid, 256, 0, st>>>(A); else nc_reduce_kernel<0><<<grid, 256, 0, st>>>(A);
  DM_CHECK_LAUNCH();
  if (splits > 1) {
    const long long total = (long long)groups * A.C;
    if (splits <= 8)
      reduce_finalize_small_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(g_ws, splits, total, A.C, A.scale, A.out1, A.out2, 0);
    else
      reduce_finalize_kernel<<<finalize_grid(total), 1024, 0, st>>>(g_ws, splits, groups, A.C, A.scale, A.out1, A.out2, 0);
    DM_CHECK_LAUNCH();
  }
  return DM_OK;
}

// db[c] += sum_p dy[p][c]: the bias gradient of a convolution that is not followed by a BatchNorm.
// Same thread mapping as the BatchNorm kernels (a thread owns 8 channels, 4 independent 16-byte loads in
// flight), block-level reduction in shared memory, per-block partial rows added by reduce_finalize_kernel.
__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ dy, int lddy, float* __restrict__ db, unsigned P,
                                                      int C, int VPB, int R) {
  extern __shared__ float sm[];          // [threads][8]
  const int cvl = threadIdx.x % VPB, r = threadIdx.x / VPB;
  const int c0 = (blockIdx.y * VPB + cvl) * 8;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (c0 < C) {
    const unsigned step = gridDim.x * R;
    unsigned p = blockIdx.x * R + r;
    uint4 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned pu = p + u * step;
      w[u] = pu < P ? dm::ldg16(dy + (long long)pu * lddy + c0) : make_uint4(0u, 0u, 0u, 0u);
    }
    while (p < P) {                                   // software pipeline, see bn_stats_kernel
      const unsigned pn = p + 4 * step;
      uint4 wn[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned pu = pn + u * step;
        wn[u] = pu < P ? dm::ldg16(dy + (long long)pu * lddy + c0) : make_uint4(0u, 0u, 0u, 0u);
      }
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) { dm::unpack8(w[u], v[u]); w[u] = wn[u]; }
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += (v[0][j] + v[1][j]) + (v[2][j] + v[3][j]);
      p = pn;
    }
  }
  float* mine = sm + threadIdx.x * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = s[j];
  __syncthreads();
  if (r == 0 && c0 < C) {
    for (int rr = 1; rr < R; ++rr) {
      const float* o = sm + (rr * VPB + cvl) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += o[j];
    }
    float* row = db + (long long)blockIdx.x * 2 * C;          // db = partial workspace here: [blocks][2][C], second half unused
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < C) row[c0 + j] = s[j];
  }
}

// Train-mode BatchNorm statistics of a stored (bf16) convolution output: per-block partial sums of y and
// y^2, written without atomics to part[blockIdx.x][2][ld] -- the layout dm_bn_finalize consumes.  Fixed
// grid + fixed summation order => bit-reproducible statistics (the conv-epilogue variant accumulates
// through shared-memory atomics and costs ~20% of the GEMM's time in shuffles).
// blockIdx.z = sample slice (GroupNorm: P pixels per sample, partial rows part[z][blockIdx.x]; BatchNorm: gridDim.z == 1).
__global__ void __launch_bounds__(256) bn_stats_kernel(const bf16* __restrict__ y, int ldy, float* __restrict__ part, int ld,
                                                        unsigned P, int C, int VPB, int R) {
  extern __shared__ float sm[];          // [threads][16]
  const int cvl = threadIdx.x % VPB, r = threadIdx.x / VPB;
  const int c0 = (blockIdx.y * VPB + cvl) * 8;
  y += (long long)blockIdx.z * P * ldy;
  part += (long long)blockIdx.z * gridDim.x * 2 * ld;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (c0 < C) {
    // software pipeline: the next four pixels' loads are issued before this round's math, so eight 16-byte loads
    // per thread are in flight (ptxas otherwise interleaves each load with the previous one's consumers)
    const unsigned step = gridDim.x * R;
    unsigned p = blockIdx.x * R + r;
    uint4 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const unsigned pu = p + u * step;
      w[u] = pu < P ? dm::ldg16(y + (long long)pu * ldy + c0) : make_uint4(0u, 0u, 0u, 0u);
    }
    while (p < P) {
      const unsigned pn = p + 4 * step;
      uint4 wn[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned pu = pn + u * step;
        wn[u] = pu < P ? dm::ldg16(y + (long long)pu * ldy + c0) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
        dm::unpack8(w[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1[j] += v[j]; s2[j] = fmaf(v[j], v[j], s2[j]); }
        w[u] = wn[u];
      }
      p = pn;
    }
  }
  float* mine = sm + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mine[j] = s1[j]; mine[8 + j] = s2[j]; }
  __syncthreads();
  if (r == 0 && c0 < C) {
    for (int rr = 1; rr < R; ++rr) {
      const float* o = sm + (rr * VPB + cvl) * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1[j] += o[j]; s2[j] += o[8 + j]; }
    }
    float* g = part + (long long)blockIdx.x * 2 * ld;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < C) { g[c0 + j] = s1[j]; g[ld + c0 + j] = s2[j]; }
  }
}

// ---------------------------------------------------------------------------------- BatchNorm
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* partials, int m_tiles, int ld, int C, double count,
                                                            float* mean, float* invstd, float* rmean, float* rvar,
                                                            float momentum, float eps, const float* conv_bias) {
  // 16 channels x 64 row slices per block, eight partial rows' loads in flight per thread: the kernel is a chain of
  // dependent round trips to L2/DRAM otherwise (13 us for 592 rows in the 32-slice, 4-row version)
  __shared__ double s1[64][17], s2[64][17];
  const int cl = threadIdx.x & 15, rr = threadIdx.x >> 4;
  const int c = blockIdx.x * 16 + cl;
  double a = 0.0, b = 0.0;
  float o_rm = 0.f, o_rv = 0.f, cb = 0.f;         // the closing thread's operands, requested before the partial rows
  if (rr == 0 && c < C && rmean) { o_rm = __ldcg(rmean + c); o_rv = __ldcg(rvar + c); cb = conv_bias ? __ldg(conv_bias + c) : 0.f; }
  if (c < C)
    for (int t = rr; t < m_tiles; t += 512) {
      float x[8], y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int tt = t + 64 * u;
        const bool ok = tt < m_tiles;
        const float* row = partials + (long long)(ok ? tt : t) * 2 * ld;
        x[u] = ok ? __ldg(row + c) : 0.f; y[u] = ok ? __ldg(row + ld + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { a += (double)x[u]; b += (double)y[u]; }
    }
  s1[rr][cl] = a; s2[rr][cl] = b;
  __syncthreads();
  if (rr < 8) {                                   // two-level fixed-order combine: 8 slices each, then 8 partials
    a = 0.0; b = 0.0;
    for (int k = 0; k < 8; ++k) { a += s1[rr * 8 + k][cl]; b += s2[rr * 8 + k][cl]; }
  }
  __syncthreads();
  if (rr < 8) { s1[rr][cl] = a; s2[rr][cl] = b; }
  __syncthreads();
  if (rr == 0 && c < C) {
    if (m_tiles > 0) {
      for (int k = 1; k < 8; ++k) { a += s1[k][cl]; b += s2[k][cl]; }
      const double mu = a / count;
      double var = b / count - mu * mu;
      if (var < 0.0) var = 0.0;
      mean[c] = (float)mu;
      invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
      if (rmean) {
        const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
        // conv_bias: the producing conv left its bias out of y (the batch norm cancels it); the tracked mean is of conv + bias
        rmean[c] = (1.f - momentum) * o_rm + momentum * ((float)mu + cb);
        rvar[c] = (1.f - momentum) * o_rv + momentum * (float)unb;
      }
    } else {
      mean[c] = o_rm;
      invstd[c] = 1.0f / sqrtf(o_rv + eps);
    }
  }
}

// z = act(y * sc + sh), sc = invstd*gamma, sh = beta - mean*sc.
// The streaming BatchNorm kernels are issue-bound on the GELU, not on HBM, so their inner loops run on packed
// fp32 pairs (dm::f32x2: FFMA2/FMUL2, see common.cuh): ~11 issue slots per element forward and ~14 backward
// instead of 24-30.  A thread owns 4 channels (8-byte vectors, two pairs) for its whole life -- coefficients in
// registers -- and keeps four pixels' loads in flight.
using dm::f32x2; using dm::pk2; using dm::bc2; using dm::upk2; using dm::fma2; using dm::mul2; using dm::add2;
using dm::bf2_to_f2; using dm::f2_to_bf2;

__device__ __forceinline__ f32x2 coef_pair(const float* __restrict__ p, int c, int C) {
  return pk2(c < C ? __ldg(p + c) : 0.f, c + 1 < C ? __ldg(p + c + 1) : 0.f);
}
// Per-channel affine u = A2*y + B2 in front of the activation.  BatchNorm (rows == 0): A2 = invstd*gamma, B2 = beta - mean*A2
// from the per-channel vectors.  GroupNorm (rows != 0): ``invstd`` / ``mean`` point at the per-(sample, channel) rows
// sc[n][c] / sh[n][c] that gn_fold_fwd_kernel prepared, sample n = blockIdx.z.
__device__ __forceinline__ void norm_affine(const float* __restrict__ mean, const float* __restrict__ invstd,
                                            const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int c0, int C,
                                            f32x2 (&A2)[4], f32x2 (&B2)[4]) {
  if (rows) {
    const long long o = (long long)blockIdx.z * C;
#pragma unroll
    for (int j = 0; j < 4; ++j) { A2[j] = coef_pair(invstd + o, c0 + 2 * j, C); B2[j] = coef_pair(mean + o, c0 + 2 * j, C); }
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + 2 * j;
    A2[j] = mul2(coef_pair(invstd, c, C), coef_pair(gamma, c, C));
    B2[j] = fma2(mul2(coef_pair(mean, c, C), bc2(-1.0f)), A2[j], coef_pair(beta, c, C));
  }
}
// Streaming skeleton shared by the three kernels below: thread (cvl, r) of block b owns the 8 channels c0..c0+7
// (one 16-byte vector; pad lanes carry zero coefficients, so they stay zero) of the pixels p0, p0+step, ...;
// pointers advance by a constant byte stride, rounds of U pixels run without bounds checks except on the
// prefetch of the following round, whose loads are issued before this round's math.
struct PixRun { unsigned p0, n, step; };
constexpr int kRingS = 3, kRingU = 2;      // cp.async ring of the two-operand kernels: 3 rounds x 2 pixels x 2 operands x 16 B per thread
__device__ __forceinline__ PixRun pix_run(unsigned P, int R, int r) {
  PixRun q;
  q.step = gridDim.x * R; q.p0 = blockIdx.x * R + r;
  q.n = q.p0 < P ? (P - q.p0 + q.step - 1) / q.step : 0u;
  return q;
}
__device__ __forceinline__ uint4 zero4() { return make_uint4(0u, 0u, 0u, 0u); }

template <int ACT>
__global__ void __launch_bounds__(256) bn_fwd_kernel(const bf16* __restrict__ y, int ldy, const float* __restrict__ mean,
                                                      const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, bf16* __restrict__ z, int ldz,
                                                      unsigned P, int C, int VPB, int R, int rows) {
  const int cvl = threadIdx.x % VPB, r = threadIdx.x / VPB;
  const int c0 = (blockIdx.y * VPB + cvl) * 8;
  PixRun q = pix_run(P, R, r);
  if (c0 >= C || q.n == 0) return;
  y += (long long)blockIdx.z * P * ldy; z += (long long)blockIdx.z * P * ldz;          // sample slice (GroupNorm)
  f32x2 sc[4], sh[4];
  norm_affine(mean, invstd, gamma, beta, rows, c0, C, sc, sh);
  constexpr int U = 4;
  const char* src = reinterpret_cast<const char*>(y + (long long)q.p0 * ldy + c0);
  char* dst = reinterpret_cast<char*>(z + (long long)q.p0 * ldz + c0);
  const long long ss = (long long)q.step * ldy * 2, ds = (long long)q.step * ldz * 2;
  auto emit = [&](const uint4& w, char* o) {
    uint4 v;
    v.x = f2_to_bf2(dm::act2<ACT>(fma2(bf2_to_f2(w.x), sc[0], sh[0])));
    v.y = f2_to_bf2(dm::act2<ACT>(fma2(bf2_to_f2(w.y), sc[1], sh[1])));
    v.z = f2_to_bf2(dm::act2<ACT>(fma2(bf2_to_f2(w.z), sc[2], sh[2])));
    v.w = f2_to_bf2(dm::act2<ACT>(fma2(bf2_to_f2(w.w), sc[3], sh[3])));
    *reinterpret_cast<uint4*>(o) = v;
  };
  uint4 w[U];
#pragma unroll
  for (int u = 0; u < U; ++u) w[u] = (unsigned)u < q.n ? dm::ldg16(src + u * ss) : zero4();
  unsigned n = q.n;
  while (n >= (unsigned)U) {
    src += U * ss;
    const unsigned rem = n - U;
    uint4 wn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) wn[u] = (unsigned)u < rem ? dm::ldg16(src + u * ss) : zero4();
#pragma unroll
    for (int u = 0; u < U; ++u) { emit(w[u], dst + u * ds); w[u] = wn[u]; }
    dst += U * ds;
    n = rem;
  }
#pragma unroll
  for (int u = 0; u < U - 1; ++u)
    if ((unsigned)u < n) emit(w[u], dst + u * ds);
}

// BatchNorm + activation backward.  With xhat = a*y + b (a = invstd, b = -mean*invstd) and the
// pre-activation u = A2*y + B2 (A2 = a*gamma, B2 = b*gamma + beta):
//   g  = dz * act'(u)
//   dy = gamma*invstd * (g - mean_p(g) - xhat*mean_p(g*xhat))  =  k0*g - K2*y - K1
// so the streaming passes need only (A2, B2) resp. (A2, B2, k0, K1, K2) per channel; the sums over
// xhat are recovered from sums over y in the finalize step (double precision).
//
// pass 1: per-block partial sums of g and g*y, written (no atomics, no zeroing) to part[block][2][C].
// (sum y is not needed: in training mode it is count*mean by construction, so sum xhat = 0 and the gradient of a
// convolution bias feeding a batch-statistics norm is exactly zero.)
template <int ACT>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16* __restrict__ dz, int lddz, const bf16* __restrict__ y,
                                                             int ldy, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ part,
                                                             unsigned P, int C, int VPB, int R, int rows) {
  extern __shared__ float sm[];          // [threads][16]
  const int cvl = threadIdx.x % VPB, r = threadIdx.x / VPB;
  const int c0 = (blockIdx.y * VPB + cvl) * 8;
  dz += (long long)blockIdx.z * P * lddz; y += (long long)blockIdx.z * P * ldy;        // sample slice (GroupNorm)
  part += (long long)blockIdx.z * gridDim.x * 2 * C;
  f32x2 s1[4], s2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) s1[j] = s2[j] = bc2(0.f);
  PixRun q = pix_run(P, R, r);
  if (c0 < C && q.n > 0) {
    f32x2 A2[4], B2[4];
    norm_affine(mean, invstd, gamma, beta, rows, c0, C, A2, B2);
    const char* pg = reinterpret_cast<const char*>(dz + (long long)q.p0 * lddz + c0);
    const char* py = reinterpret_cast<const char*>(y + (long long)q.p0 * ldy + c0);
    const long long sg = (long long)q.step * lddz * 2, sy = (long long)q.step * ldy * 2;
    auto accum = [&](const uint4& g, const uint4& v) {     // dz = 0 (tail padding) contributes nothing
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, vw[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const f32x2 yy = bf2_to_f2(vw[j]);
        const f32x2 gg = mul2(bf2_to_f2(gw[j]), dm::act_grad2<ACT>(fma2(yy, A2[j], B2[j])));
        s1[j] = add2(s1[j], gg); s2[j] = fma2(gg, yy, s2[j]);
      }
    };
    // ring of kRingS rounds x kRingU pixels x 2 operands, 16 bytes per (slot, thread)
    const uint32_t ring = dm::smem_u32(sm) + threadIdx.x * 16u, slot = blockDim.x * 16u;
    const unsigned rounds = (q.n + kRingU - 1) / kRingU;
    auto issue = [&](unsigned rd) {
      const uint32_t base = ring + (rd % kRingS) * (kRingU * 2) * slot;
#pragma unroll
      for (int u = 0; u < kRingU; ++u) {
        const unsigned k = rd * kRingU + u;
        const bool ok = k < q.n;
        const long long kk = ok ? (long long)k : 0;
        dm::cp_async16(base + (2 * u) * slot, pg + kk * sg, ok);
        dm::cp_async16(base + (2 * u + 1) * slot, py + kk * sy, ok);
      }
    };
#pragma unroll
    for (int s = 0; s < kRingS - 1; ++s) { if ((unsigned)s < rounds) issue(s); dm::cp_async_commit(); }
    for (unsigned rd = 0; rd < rounds; ++rd) {
      if (rd + kRingS - 1 < rounds) issue(rd + kRingS - 1);
      dm::cp_async_commit();
      dm::cp_async_wait<kRingS - 1>();
      const uint32_t base = ring + (rd % kRingS) * (kRingU * 2) * slot;
#pragma unroll
      for (int u = 0; u < kRingU; ++u) accum(dm::lds16(base + (2 * u) * slot), dm::lds16(base + (2 * u + 1) * slot));
    }
    dm::cp_async_wait<0>();
  }
  __syncthreads();                       // the ring is reused for the block reduction
  float s[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) { upk2(s1[j], s[2 * j], s[2 * j + 1]); upk2(s2[j], s[8 + 2 * j], s[8 + 2 * j + 1]); }
  float* mine = sm + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 16; ++j) mine[j] = s[j];
  __syncthreads();
  if (r == 0 && c0 < C) {
    for (int rr = 1; rr < R; ++rr) {
      const float* o = sm + (rr * VPB + cvl) * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) s[j] += o[j];
    }
    float* g = part + (long long)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < C) { g[c0 + j] = s[j]; g[C + c0 + j] = s[8 + j]; }
  }
}

// pass 2: block partials -> per-channel coefficients coef[3][C] = (k0, K1, K2) of the apply pass;
// dgamma += sum g*xhat, dbeta += sum g, and (eval mode) the bias gradient of the convolution feeding this norm.
// 32 channels x 32 row slices per block; each thread issues its partial-row loads four at a time (a plain
// dependent loop costs one L2 round trip per row: 13 us for 600 rows).
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const float* __restrict__ part, int nblk, int C, double P,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, float* __restrict__ coef,
                                                                float* dgamma, float* dbeta, float* dbias, int training) {
  __shared__ float sh[2][64][17];
  const int cl = threadIdx.x & 15, rr = threadIdx.x >> 4;          // 16 channels x 64 row slices, see bn_finalize_kernel
  const int c = blockIdx.x * 16 + cl;
  float f0 = 0.f, f1 = 0.f;
  // the closing thread's operands (and the old values of the accumulate-into-.grad outputs) are requested up front:
  // read at the end they are three more dependent round trips behind the partial-row loads
  const bool closer = rr == 0 && c < C;
  float p_inv = 0.f, p_mean = 0.f, p_gamma = 0.f, o_beta = 0.f, o_gamma = 0.f, o_bias = 0.f;
  if (closer) {
    p_inv = __ldg(invstd + c); p_mean = __ldg(mean + c); p_gamma = __ldg(gamma + c);
    o_beta = __ldcg(dbeta + c); o_gamma = __ldcg(dgamma + c);
    if (dbias && !training) o_bias = __ldcg(dbias + c);
  }
  if (c < C) {
    for (int t = rr; t < nblk; t += 512) {
      float x0[8], x1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int tt = t + 64 * u;
        const bool ok = tt < nblk;
        const float* g = part + (long long)(ok ? tt : t) * 2 * C;
        x0[u] = ok ? __ldg(g + c) : 0.f; x1[u] = ok ? __ldg(g + C + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { f0 += x0[u]; f1 += x1[u]; }
    }
  }
  sh[0][rr][cl] = f0; sh[1][rr][cl] = f1;
  __syncthreads();
  if (closer) {
    double sg = 0.0, sgy = 0.0;
    for (int k = 0; k < 64; ++k) { sg += (double)sh[0][k][cl]; sgy += (double)sh[1][k][cl]; }
    const double a = (double)p_inv, b = -(double)p_mean * a, ga = (double)p_gamma;
    const double sgx = a * sgy + b * sg;          // sum g*xhat
    const double k0 = ga * a;
    const double m1 = training ? sg / P : 0.0, m2 = training ? sgx / P : 0.0;
    coef[c] = (float)k0;
    coef[C + c] = (float)(-k0 * (m1 + m2 * b));   // -K1
    coef[2 * C + c] = (float)(-k0 * m2 * a);      // -K2
    dbeta[c] = o_beta + (float)sg;
    dgamma[c] = o_gamma + (float)sgx;
    if (dbias && !training) dbias[c] = o_bias + (float)(k0 * sg);
  }
}

// pass 3: dy = k0*g - K2*y - K1   (coef rows hold k0, -K1, -K2)
template <int ACT>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const bf16* __restrict__ dz, int lddz, const bf16* __restrict__ y,
                                                            int ldy, const float* __restrict__ mean,
                                                            const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ coef,
                                                            bf16* __restrict__ dy, int lddy, unsigned P, int C, int VPB, int R,
                                                            int rows) {
  const int cvl = threadIdx.x % VPB, r = threadIdx.x / VPB;
  const int c0 = (blockIdx.y * VPB + cvl) * 8;
  PixRun q = pix_run(P, R, r);
  if (c0 >= C || q.n == 0) return;
  dz += (long long)blockIdx.z * P * lddz; y += (long long)blockIdx.z * P * ldy; dy += (long long)blockIdx.z * P * lddy;
  if (rows) coef += (long long)blockIdx.z * 3 * C;                                       // coef[n][3][C] (GroupNorm)
  f32x2 A2[4], B2[4], k0[4], nK1[4], nK2[4];
  norm_affine(mean, invstd, gamma, beta, rows, c0, C, A2, B2);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + 2 * j;
    k0[j] = coef_pair(coef, c, C); nK1[j] = coef_pair(coef + C, c, C); nK2[j] = coef_pair(coef + 2 * C, c, C);
  }
  extern __shared__ float sm[];
  const char* pg = reinterpret_cast<const char*>(dz + (long long)q.p0 * lddz + c0);
  const char* py = reinterpret_cast<const char*>(y + (long long)q.p0 * ldy + c0);
  char* dst = reinterpret_cast<char*>(dy + (long long)q.p0 * lddy + c0);
  const long long sg = (long long)q.step * lddz * 2, sy = (long long)q.step * ldy * 2, ds = (long long)q.step * lddy * 2;
  auto emit = [&](const uint4& g, const uint4& v, char* o) {
    const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, vw[4] = {v.x, v.y, v.z, v.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const f32x2 yy = bf2_to_f2(vw[j]);
      const f32x2 gg = mul2(bf2_to_f2(gw[j]), dm::act_grad2<ACT>(fma2(yy, A2[j], B2[j])));
      ow[j] = f2_to_bf2(fma2(k0[j], gg, fma2(nK2[j], yy, nK1[j])));
    }
    *reinterpret_cast<uint4*>(o) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  };
  const uint32_t ring = dm::smem_u32(sm) + threadIdx.x * 16u, slot = blockDim.x * 16u;
  const unsigned rounds = (q.n + kRingU - 1) / kRingU;
  auto issue = [&](unsigned rd) {
    const uint32_t base = ring + (rd % kRingS) * (kRingU * 2) * slot;
#pragma unroll
    for (int u = 0; u < kRingU; ++u) {
      const unsigned k = rd * kRingU + u;
      const bool ok = k < q.n;
      const long long kk = ok ? (long long)k : 0;
      dm::cp_async16(base + (2 * u) * slot, pg + kk * sg, ok);
      dm::cp_async16(base + (2 * u + 1) * slot, py + kk * sy, ok);
    }
  };
#pragma unroll
  for (int s = 0; s < kRingS - 1; ++s) { if ((unsigned)s < rounds) issue(s); dm::cp_async_commit(); }
  for (unsigned rd = 0; rd < rounds; ++rd) {
    if (rd + kRingS - 1 < rounds) issue(rd + kRingS - 1);
    dm::cp_async_commit();
    dm::cp_async_wait<kRingS - 1>();
    const uint32_t base = ring + (rd % kRingS) * (kRingU * 2) * slot;
#pragma unroll
    for (int u = 0; u < kRingU; ++u) {
      const unsigned k = rd * kRingU + u;
      if (k < q.n) emit(dm::lds16(base + (2 * u) * slot), dm::lds16(base + (2 * u + 1) * slot), dst + (long long)k * ds);
    }
  }
  dm::cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------- GroupNorm
// The streaming passes ARE the BatchNorm kernels (bn_stats_kernel, bn_fwd_kernel, bn_bwd_reduce_kernel, bn_bwd_apply_kernel)
// launched with one grid.z slice per sample and ``rows`` = 1: the statistics of a (sample, group) fold into per-(sample,
// channel) rows scale/shift[n][c] once (gn_fold_fwd_kernel), the backward sums into rows (k0, -K1, -K2)[n][c]
// (gn_fold_bwd_kernel), so the big passes are load -> fma -> activation -> store with the coefficients in registers.
//
// pass 2 (forward): one block per (sample, group): statistics -> mean/rstd and the coefficient rows
// sc[n][c] = rstd*gamma, sh[n][c] = beta - mean*rstd*gamma
__global__ void __launch_bounds__(256) gn_fold_fwd_kernel(const float* __restrict__ part, int nblk, int C, int G, double cnt,
                                                           float eps, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* mean, float* rstd,
                                                           float* sc, float* sh, float* sx) {
  __shared__ double red[2][256];
  const int n = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < cg * nblk; i += blockDim.x) {
    const int c = g * cg + i % cg, t = i / cg;
    const float* row = part + ((long long)n * nblk + t) * 2 * C;
    a += (double)row[c]; b += (double)row[C + c];
  }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
    __syncthreads();
  }
  const double mu = red[0][0] / cnt;
  double var = red[1][0] / cnt - mu * mu; if (var < 0) var = 0;
  const double rs = 1.0 / sqrt(var + (double)eps);
  if (threadIdx.x == 0) { mean[blockIdx.x] = (float)mu; rstd[blockIdx.x] = (float)rs; }
  for (int i = threadIdx.x; i < cg; i += blockDim.x) {
    const int c = g * cg + i;
    const double k = rs * (double)gamma[c];
    sc[(long long)n * C + c] = (float)k;
    sh[(long long)n * C + c] = (float)((double)beta[c] - mu * k);
  }
  (void)sx;
}
// backward pass 2: one block per (sample, group).  With Sg = sum g, Sgx = sum g*x per channel:
//   m1 = sum_c gamma*Sg / cnt,  m2 = sum_c gamma*rstd*(Sgx - mean*Sg) / cnt
//   dx = (rstd*gamma)*g - (rstd^2*m2)*x - (rstd*m1 - rstd^2*m2*mean)        -> coef[n][3][C] = (k0, -K1, -K2)
//   dgamma[c] += rstd*(Sgx - mean*Sg), dbeta[c] += Sg      (atomic across samples)
__global__ void __launch_bounds__(256) gn_fold_bwd_kernel(const float* __restrict__ part, int nblk, int C, int G, double cnt,
                                                           const float* __restrict__ gamma, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ coef,
                                                           float* dgamma, float* dbeta) {
  __shared__ double red[2][256];
  const int n = blockIdx.x / G, g = blockIdx.x % G, cg = C / G;
  const double mu = (double)mean[blockIdx.x], rs = (double)rstd[blockIdx.x];
  // per-channel sums over the partial rows: (channel, row-slice) threads, then a fixed-order combine
  __shared__ float ps[2][256];
  const int slices = cg < 256 ? 256 / cg : 1;
  double a = 0.0, b = 0.0;
  for (int i0 = 0; i0 < cg; i0 += 256) {
    const int i = i0 + (int)threadIdx.x % (cg < 256 ? cg : 256), sl = (int)threadIdx.x / (cg < 256 ? cg : 256);
    float f0 = 0.f, f1 = 0.f;
    if (i < cg && sl < slices) {
      const int c = g * cg + i;
      for (int q = sl; q < nblk; q += slices) {
        const float* row = part + ((long long)n * nblk + q) * 2 * C;
        f0 += row[c]; f1 += row[C + c];
      }
    }
    ps[0][threadIdx.x] = f0; ps[1][threadIdx.x] = f1;
    __syncthreads();
    if (sl == 0 && i < cg) {
      const int c = g * cg + i, stride = cg < 256 ? cg : 256;
      double sg = 0.0, sgx = 0.0;
      for (int k = 0; k < slices; ++k) { sg += (double)ps[0][k * stride + (i - i0)]; sgx += (double)ps[1][k * stride + (i - i0)]; }
      const double ga = (double)gamma[c];
      a += ga * sg; b += ga * rs * (sgx - mu * sg);
      atomicAdd(dbeta + c, (float)sg);
      atomicAdd(dgamma + c, (float)(rs * (sgx - mu * sg)));
    }
    __syncthreads();
  }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red[0][threadIdx.x] += red[0][threadIdx.x + o]; red[1][threadIdx.x] += red[1][threadIdx.x + o]; }
    __syncthreads();
  }
  const double m1 = red[0][0] / cnt, m2 = red[1][0] / cnt;
  for (int i = threadIdx.x; i < cg; i += blockDim.x) {
    const int c = g * cg + i;
    float* row = coef + (long long)n * 3 * C;
    row[c] = (float)(rs * (double)gamma[c]);                 // k0
    row[C + c] = (float)-(rs * m1 - rs * rs * m2 * mu);       // -K1   (bn_bwd_apply_kernel computes k0*g + (-K2)*x + (-K1))
    row[2 * C + c] = (float)-(rs * rs * m2);                  // -K2
  }
}

// ---------------------------------------------------------------------------------- SE / residual
struct SeFwd {
  const bf16* x2; int ld2; const float* gate; const bf16* res; int ldr; bf16* out; int ldo; int C, HW; float scale;
  __device__ void operator()(unsigned p, int c0) const {
    float a[8], r[8], gt[8];
    load8(x2 + (long long)p * ld2 + c0, a);
    load8(res + (long long)p * ldr + c0, r);
    const int n = p / HW;
    if (gate) ldp8(gate + (long long)n * C, c0, C, gt);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (c0 + j < C) ? (r[j] + a[j] * (gate ? gt[j] : 1.f)) * scale : 0.f;
    store8(out + (long long)p * ldo + c0, a);
  }
};
struct SeBwd {
  const bf16* dout; int lddo; const float *gate, *dpool; bf16* dx2; int lddx2; bf16* dres; int lddr; int C, HW; float scale, invHW;
  __device__ void operator()(unsigned p, int c0) const {
    float g[8], gt[8], dp[8], o[8];
    load8(dout + (long long)p * lddo + c0, g);
    const int n = p / HW;
    if (gate) ldp8(gate + (long long)n * C, c0, C, gt);
    if (dpool) ldp8(dpool + (long long)n * C, c0, C, dp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gs = g[j] * scale;
      o[j] = (c0 + j < C) ? gs * (gate ? gt[j] : 1.f) + (dpool ? dp[j] * invHW : 0.f) : 0.f;
      g[j] = (c0 + j < C) ? gs : 0.f;
    }
    store8(dx2 + (long long)p * lddx2 + c0, o);
    if (dres) store8(dres + (long long)p * lddr + c0, g);
  }
};

// ---------------------------------------------------------------------------------- CoordAttn gates
struct CaFwd {
  const bf16* x; int ldx; const float *ah, *aw; bf16* out; int ldo; int C, H, W;
  __device__ void operator()(unsigned p, int c0) const {
    float v[8], a[8], b[8];
    load8(x + (long long)p * ldx + c0, v);
    const int w = p % W, nh = p / W;            // nh = n*H + h
    const int n = nh / H;
    ldp8(ah + (long long)nh * C, c0, C, a);
    ldp8(aw + ((long long)n * W + w) * C, c0, C, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? v[j] * (a[j] + b[j]) : 0.f;
    store8(out + (long long)p * ldo + c0, v);
  }
};
struct CaBwd {
  const bf16* dout; int lddo; const float *ah, *aw, *dxh, *dxw; bf16* dx; int lddx; int C, H, W; float invW, invH;
  __device__ void operator()(unsigned p, int c0) const {
    float g[8], a[8], b[8], e[8], f[8];
    load8(dout + (long long)p * lddo + c0, g);
    const int w = p % W, nh = p / W;
    const int n = nh / H;
    ldp8(ah + (long long)nh * C, c0, C, a);
    ldp8(aw + ((long long)n * W + w) * C, c0, C, b);
    ldp8(dxh + (long long)nh * C, c0, C, e);
    ldp8(dxw + ((long long)n * W + w) * C, c0, C, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = (c0 + j < C) ? g[j] * (a[j] + b[j]) + e[j] * invW + f[j] * invH : 0.f;
    store8(dx + (long long)p * lddx + c0, g);
  }
};

// ---------------------------------------------------------------------------------- FiLM
struct FilmFwd {
  const bf16* x; int ldx; const float *ce, *te; bf16* out; int ldo; int C, HW;
  __device__ void operator()(unsigned p, int c0) const {
    float v[8], a[8], b[8];
    load8(x + (long long)p * ldx + c0, v);
    const int n = p / HW;
    ldp8(ce + (long long)n * C, c0, C, a); ldp8(te + (long long)n * C, c0, C, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? a[j] * v[j] + b[j] : 0.f;
    store8(out + (long long)p * ldo + c0, v);
  }
};
struct ScaleNC {   // dx = dout * ce[n,c]
  const bf16* dout; int lddo; const float* ce; bf16* dx; int lddx; int C, HW;
  __device__ void operator()(unsigned p, int c0) const {
    float v[8], a[8];
    load8(dout + (long long)p * lddo + c0, v);
    ldp8(ce + (long long)(p / HW) * C, c0, C, a);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c0 + j < C) ? v[j] * a[j] : 0.f;
    store8(dx + (long long)p * lddx + c0, v);
  }
};

// ---------------------------------------------------------------------------------- upsample + cat
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_src(int o, int in, float scale) {
  // at::native upsample_bilinear2d, align_corners=True: src = scale*dst, scale = (in-1)/(out-1)
  const float s = scale * (float)o;
  Lerp L; L.i0 = (int)s; L.i1 = L.i0 + (L.i0 < in - 1 ? 1 : 0); L.l1 = s - (float)L.i0; L.l0 = 1.0f - L.l1;
  return L;
}
// Bilinear blend of one bf16x2 word from each of the four source pixels on packed fp32 pairs (FMUL2 / FFMA2: half the
// issue slots of the scalar form -- the upsample kernels are issue-bound, not HBM-bound).  One fixed rounding sequence,
// shared by the per-pixel and the quad kernel, so the two stay bit-identical:
//   top = fma(x1, v01, x0*v00), bot = fma(x1, v11, x0*v10), out = fma(y1, bot, y0*top)
__device__ __forceinline__ f32x2 xlerp2(uint32_t v0, uint32_t v1, f32x2 x0, f32x2 x1) {
  return fma2(x1, bf2_to_f2(v1), mul2(x0, bf2_to_f2(v0)));
}
__device__ __forceinline__ uint32_t ylerp2_bf(f32x2 top, f32x2 bot, f32x2 y0, f32x2 y1) {
  return f2_to_bf2(fma2(y1, bot, mul2(y0, top)));
}
// zero the lanes of a stored 8-channel vector that lie past the tensor's channel count (ragged last vector only)
__device__ __forceinline__ void mask_tail(uint4& o, int valid) {
  if (valid >= 8) return;
  uint32_t* w = &o.x;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (2 * q >= valid) w[q] = 0u;
    else if (2 * q + 1 >= valid) w[q] &= 0x0000ffffu;
  }
}
struct UpcatFwd {
  const bf16* a; int lda, Ca; const bf16* b; int ldb, Cb; bf16* out; int ldo; int h, w; float sy, sx;
  int nb;          // samples held by b: sample n of the batch reads b[n % nb] (a skip tensor shared by the CFG halves); 0 = N
  __device__ void operator()(unsigned p, int c0) const {
    const int W2 = 2 * w, H2 = 2 * h;
    const int ox = p % W2, oy = (p / W2) % H2, n = p / (W2 * H2);
    const Lerp Y = lerp_src(oy, h, sy), X = lerp_src(ox, w, sx);
    const bf16* src; int ld, c, C, ns = n;
    if (c0 < Ca) { src = a; ld = lda; c = c0; C = Ca; } else { src = b; ld = ldb; c = c0 - Ca; C = Cb; if (nb > 0) ns = n % nb; }
    const long long base = (long long)ns * h * w;
    const uint4 r00 = dm::ldg16(src + (base + (long long)Y.i0 * w + X.i0) * ld + c);
    const uint4 r01 = dm::ldg16(src + (base + (long long)Y.i0 * w + X.i1) * ld + c);
    const uint4 r10 = dm::ldg16(src + (base + (long long)Y.i1 * w + X.i0) * ld + c);
    const uint4 r11 = dm::ldg16(src + (base + (long long)Y.i1 * w + X.i1) * ld + c);
    const f32x2 x0 = bc2(X.l0), x1 = bc2(X.l1), y0 = bc2(Y.l0), y1 = bc2(Y.l1);
    const uint32_t *p00 = &r00.x, *p01 = &r01.x, *p10 = &r10.x, *p11 = &r11.x;
    uint4 o;
    uint32_t* po = &o.x;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      po[q] = ylerp2_bf(xlerp2(p00[q], p01[q], x0, x1), xlerp2(p10[q], p11[q], x0, x1), y0, y1);
    mask_tail(o, C - c);
    *reinterpret_cast<uint4*>(out + (long long)p * ldo + c0) = o;
  }
};
// gather-form backward: every input pixel sums the output pixels that sampled it (deterministic)
struct UpcatBwd {
  const bf16* dout; int lddo; bf16* da; int ldda, Ca; bf16* db; int lddb, Cb; int h, w; float sy, sx;
  __device__ void operator()(unsigned p, int c0) const {
    const int ix = p % w, iy = (p / w) % h, n = p / (w * h);
    const int W2 = 2 * w, H2 = 2 * h;
    int oy0 = sy > 0.f ? (int)floorf((float)(iy - 1) / sy) : 0; if (oy0 < 0) oy0 = 0;
    int oy1 = sy > 0.f ? (int)ceilf((float)(iy + 1) / sy) : H2 - 1; if (oy1 > H2 - 1) oy1 = H2 - 1;
    int ox0 = sx > 0.f ? (int)floorf((float)(ix - 1) / sx) : 0; if (ox0 < 0) ox0 = 0;
    int ox1 = sx > 0.f ? (int)ceilf((float)(ix + 1) / sx) : W2 - 1; if (ox1 > W2 - 1) ox1 = W2 - 1;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy) {
      const Lerp Y = lerp_src(oy, h, sy);
      const float wy = (Y.i0 == iy ? Y.l0 : 0.f) + (Y.i1 == iy ? Y.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox0; ox <= ox1; ++ox) {
        const Lerp X = lerp_src(ox, w, sx);
        const float wx = (X.i0 == ix ? X.l0 : 0.f) + (X.i1 == ix ? X.l1 : 0.f);
        if (wx == 0.f) continue;
        float g[8];
        load8(dout + (((long long)n * H2 + oy) * W2 + ox) * lddo + c0, g);
        const float wt = wy * wx;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wt * g[j];
      }
    }
    if (c0 < Ca) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c0 + j >= Ca) acc[j] = 0.f;
      store8(da + (long long)p * ldda + c0, acc);
    } else {
      const int c = c0 - Ca;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (c + j >= Cb) acc[j] = 0.f;
      store8(db + (long long)p * lddb + c, acc);
    }
  }
};

// Quad forms of the two passes above (same arithmetic, same accumulation order -> bit-identical results).
// With align_corners=True and an exact x2 factor, output rows 2k-1 and 2k both interpolate between source
// rows k-1 and k (lerp_src: (int)(o*(h-1)/(2h-1)) = k-1 for both), so a thread that owns the 2x2 output quad
// (rows 2ky-1..2ky, cols 2kx-1..2kx) needs exactly one 2x2 source block: one 16-byte load per 16-byte store
// instead of four.  Quads on the border have one valid row and/or column.
// grid: x over the (w + 1) * Cv (quad column, channel vector) pairs of one quad row, y = n * (h + 1) + ky -- one integer
// division per thread instead of three, and every address is one 64-bit base plus small steps.
__global__ void __launch_bounds__(kEwThreads) upcat_fwd_quad_kernel(UpcatFwd f, unsigned Cv) {
  const int h = f.h, w = f.w, W2 = 2 * w, H2 = 2 * h;
  const int n = blockIdx.y / (unsigned)(h + 1), ky = blockIdx.y - n * (h + 1);
  {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const int kx = i / Cv;
    if (kx > w) return;
    const int c0 = (int)(i - kx * Cv) * 8;
    const bool vy[2] = {ky >= 1, ky <= h - 1}, vx[2] = {kx >= 1, kx <= w - 1};
    const int oy[2] = {vy[0] ? 2 * ky - 1 : 2 * ky, vy[1] ? 2 * ky : 2 * ky - 1};
    const int ox[2] = {vx[0] ? 2 * kx - 1 : 2 * kx, vx[1] ? 2 * kx : 2 * kx - 1};
    const Lerp Y[2] = {lerp_src(oy[0], h, f.sy), lerp_src(oy[1], h, f.sy)};
    const Lerp X[2] = {lerp_src(ox[0], w, f.sx), lerp_src(ox[1], w, f.sx)};
    const bf16* src; int ld, c, C, ns = n;
    if (c0 < f.Ca) { src = f.a; ld = f.lda; c = c0; C = f.Ca; }
    else { src = f.b; ld = f.ldb; c = c0 - f.Ca; C = f.Cb; if (f.nb > 0) ns = n % f.nb; }
    const bf16* s00 = src + (((long long)ns * h + Y[0].i0) * w + X[0].i0) * ld + c;
    const int sdx = (X[0].i1 - X[0].i0) * ld;                       // 0 on the right border
    const long long sdy = (long long)(Y[0].i1 - Y[0].i0) * w * ld;  // 0 on the bottom border
    const uint4 r00 = dm::ldg16(s00);
    const uint4 r01 = dm::ldg16(s00 + sdx);
    const uint4 r10 = dm::ldg16(s00 + sdy);
    const uint4 r11 = dm::ldg16(s00 + sdy + sdx);
    bf16* o00 = f.out + (((long long)n * H2 + oy[0]) * W2 + ox[0]) * f.ldo + c0;
    const int odx = (ox[1] - ox[0]) * f.ldo;
    const long long ody = (long long)(oy[1] - oy[0]) * W2 * f.ldo;
    const uint32_t *p00 = &r00.x, *p01 = &r01.x, *p10 = &r10.x, *p11 = &r11.x;
    // horizontal blends of the two source rows for both output columns, shared by the two output rows
    f32x2 top[2][4], bot[2][4];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const f32x2 x0 = bc2(X[b].l0), x1 = bc2(X[b].l1);
#pragma unroll
      for (int q = 0; q < 4; ++q) { top[b][q] = xlerp2(p00[q], p01[q], x0, x1); bot[b][q] = xlerp2(p10[q], p11[q], x0, x1); }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (!vy[a]) continue;
      const f32x2 y0 = bc2(Y[a].l0), y1 = bc2(Y[a].l1);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (!vx[b]) continue;
        uint4 o;
        uint32_t* po = &o.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) po[q] = ylerp2_bf(top[b][q], bot[b][q], y0, y1);
        mask_tail(o, C - c);
        *reinterpret_cast<uint4*>(o00 + (a ? ody : 0) + (b ? odx : 0)) = o;
      }
    }
  }
}
// Backward: a thread owns a 2x2 block of input pixels and walks the 6x6 output window that touches it (rows
// 2*iy0-1 .. 2*iy0+4): 9 loads per input vector instead of 16, all of one window row in flight together, no
// data-dependent control flow.  Per input pixel the products are added in the same (row, column) order as UpcatBwd.
__global__ void __launch_bounds__(128) upcat_bwd_quad_kernel(UpcatBwd f, unsigned total, unsigned Cv) {
  const int h = f.h, w = f.w, W2 = 2 * w, H2 = 2 * h, hb = (h + 1) / 2, wb = (w + 1) / 2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned q = i / Cv;
    const int c0 = (int)(i - q * Cv) * 8;
    const int ix0 = 2 * (int)(q % (unsigned)wb); q /= (unsigned)wb;
    const int iy0 = 2 * (int)(q % (unsigned)hb);
    const int n = q / (unsigned)hb;
    float wy[2][6], wx[2][6];
    int ry[6], rx[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int oy = 2 * iy0 - 1 + r, ox = 2 * ix0 - 1 + r;
      const bool vy = oy >= 0 && oy < H2, vx = ox >= 0 && ox < W2;
      ry[r] = vy ? oy : 0; rx[r] = vx ? ox : 0;
      const Lerp Y = lerp_src(ry[r], h, f.sy), X = lerp_src(rx[r], w, f.sx);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        wy[a][r] = vy ? (Y.i0 == iy0 + a ? Y.l0 : 0.f) + (Y.i1 == iy0 + a ? Y.l1 : 0.f) : 0.f;
        wx[a][r] = vx ? (X.i0 == ix0 + a ? X.l0 : 0.f) + (X.i1 == ix0 + a ? X.l1 : 0.f) : 0.f;
      }
    }
    // accumulators on packed fp32 pairs (FFMA2: the same fma per element as the scalar form, half the issue slots)
    f32x2 acc[2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][b][j] = bc2(0.f);
    const bf16* gbase = f.dout + (long long)n * H2 * W2 * f.lddo + c0;
    int coff[6];
#pragma unroll
    for (int cc = 0; cc < 6; ++cc) coff[cc] = rx[cc] * f.lddo;
    // window rows are double-buffered: the next row's six loads are requested before this row's arithmetic
    uint4 raw[6], nxt[6];
    {
      const bf16* rowp = gbase + (long long)ry[0] * W2 * f.lddo;
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) raw[cc] = dm::ldg16(rowp + coff[cc]);
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r + 1 < 6) {
        const bf16* rowp = gbase + (long long)ry[r + 1] * W2 * f.lddo;
#pragma unroll
        for (int cc = 0; cc < 6; ++cc) nxt[cc] = dm::ldg16(rowp + coff[cc]);
      }
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) {
        const f32x2 g[4] = {bf2_to_f2(raw[cc].x), bf2_to_f2(raw[cc].y), bf2_to_f2(raw[cc].z), bf2_to_f2(raw[cc].w)};
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (r < 2 * a || r >= 2 * a + 4) continue;      // rows outside 2*iy-1 .. 2*iy+2 never sample input row iy
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            if (cc < 2 * b || cc >= 2 * b + 4) continue;
            const f32x2 wt = bc2(wy[a][r] * wx[b][cc]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[a][b][j] = fma2(wt, g[j], acc[a][b][j]);
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < 6; ++cc) raw[cc] = nxt[cc];
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (iy0 + a >= h) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (ix0 + b >= w) continue;
        const long long p = ((long long)n * h + iy0 + a) * w + ix0 + b;
        uint4 o;
        o.x = f2_to_bf2(acc[a][b][0]); o.y = f2_to_bf2(acc[a][b][1]); o.z = f2_to_bf2(acc[a][b][2]); o.w = f2_to_bf2(acc[a][b][3]);
        if (c0 < f.Ca) {
          mask_tail(o, f.Ca - c0);
          *reinterpret_cast<uint4*>(f.da + p * f.ldda + c0) = o;
        } else {
          const int c = c0 - f.Ca;
          mask_tail(o, f.Cb - c);
          *reinterpret_cast<uint4*>(f.db + p * f.lddb + c) = o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------- pooling
struct AvgPoolFwd {
  const bf16* x; int ldx; bf16* out; int ldo; int H, W, C, k, act;
  __device__ void operator()(unsigned p, int c0) const {
    const int Wo = W / k, Ho = H / k;
    const int ox = p % Wo, oy = (p / Wo) % Ho, n = p / (Wo * Ho);
    float acc[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        load8(x + (((long long)n * H + oy * k + dy) * W + ox * k + dx) * ldx + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    const float inv = 1.0f / (float)(k * k);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = (c0 + j < C) ? dm::act_f(acc[j] * inv, act) : 0.f;
    store8(out + (long long)p * ldo + c0, acc);
  }
};
struct AvgPoolBwd {
  const bf16* dout; int lddo; const bf16* x; int ldx; bf16* dx; int lddx; int H, W, C, k, act;
  __device__ void operator()(unsigned p, int c0) const {
    const int Wo = W / k, Ho = H / k;
    const int ox = p % Wo, oy = (p / Wo) % Ho, n = p / (Wo * Ho);
    float acc[8], v[8], g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx_ = 0; dx_ < k; ++dx_) {
        load8(x + (((long long)n * H + oy * k + dy) * W + ox * k + dx_) * ldx + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    const float inv = 1.0f / (float)(k * k);
    load8(dout + (long long)p * lddo + c0, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = (c0 + j < C) ? g[j] * dm::act_grad_f(acc[j] * inv, act) * inv : 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx_ = 0; dx_ < k; ++dx_)
        store8(dx + (((long long)n * H + oy * k + dy) * W + ox * k + dx_) * lddx + c0, g);
  }
};
struct MaxPoolFwd {
  const bf16* x; int ldx; bf16* out; int ldo; int H, W, C;
  __device__ void operator()(unsigned p, int c0) const {
    const int Wo = W / 2, Ho = H / 2;
    const int ox = p % Wo, oy = (p / Wo) % Ho, n = p / (Wo * Ho);
    float m[8], v[8];
    const long long b = (((long long)n * H + oy * 2) * W + ox * 2);
    load8(x + b * ldx + c0, m);
    const long long offs[3] = {1, (long long)W, (long long)W + 1};
    for (int t = 0; t < 3; ++t) {
      load8(x + (b + offs[t]) * ldx + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = (v[j] > m[j] || v[j] != v[j]) ? v[j] : m[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) if (c0 + j >= C) m[j] = 0.f;
    store8(out + (long long)p * ldo + c0, m);
  }
};
struct MaxPoolBwd {
  const bf16* dout; int lddo; const bf16* x; int ldx; bf16* dx; int lddx; int H, W, C;
  __device__ void operator()(unsigned p, int c0) const {
    const int Wo = W / 2, Ho = H / 2;
    const int ox = p % Wo, oy = (p / Wo) % Ho, n = p / (Wo * Ho);
    float m[8], v[4][8], g[8];
    int am[8];
    const long long b = (((long long)n * H + oy * 2) * W + ox * 2);
    const long long offs[4] = {0, 1, (long long)W, (long long)W + 1};
    for (int t = 0; t < 4; ++t) load8(x + (b + offs[t]) * ldx + c0, v[t]);
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = v[0][j]; am[j] = 0; }
    for (int t = 1; t < 4; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) if (v[t][j] > m[j] || v[t][j] != v[t][j]) { m[j] = v[t][j]; am[j] = t; }
    load8(dout + (long long)p * lddo + c0, g);
    for (int t = 0; t < 4; ++t) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (am[j] == t && c0 + j < C) ? g[j] : 0.f;
      store8(dx + (b + offs[t]) * lddx + c0, o);
    }
  }
};

// ---------------------------------------------------------------------------------- misc elementwise
struct MaskFma {
  const bf16* x; int ldx; const bf16* y; int ldy; const float* mask; float thresh; bf16* out; int ldo; int C;
  __device__ void operator()(unsigned p, int c0) const {
    float a[8], b[8];
    load8(y + (long long)p * ldy + c0, b);
    if (x) load8(x + (long long)p * ldx + c0, a);
    const float hi = (__ldg(mask + p) > thresh) ? 1.f : 0.f;     // exact fp32 compare, NaN -> 0 (new_scripy.py:173)
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (c0 + j < C) ? (x ? a[j] : 0.f) + b[j] * hi : 0.f;
    store8(out + (long long)p * ldo + c0, a);
  }
};
struct Axpby {
  const bf16* a; int lda; const bf16* b; int ldb; bf16* out; int ldo; int C; float sa, sb;
  __device__ void operator()(unsigned p, int c0) const {
    float u[8], v[8];
    load8(a + (long long)p * lda + c0, u);
    if (b) load8(b + (long long)p * ldb + c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = (c0 + j < C) ? sa * u[j] + (b ? sb * v[j] : 0.f) : 0.f;
    store8(out + (long long)p * ldo + c0, u);
  }
};
struct S2D {
  const bf16* x; int ldx; bf16* y; int ldy; int H, W, C, k, chan_major;    // x [N,H*k,W*k,C] -> y [N,H,W,k*k*C]
  __device__ void operator()(unsigned p, int c0) const {
    // p indexes INPUT pixels of x.  Output channel = tap*C + c (tap-major) or c*k*k + tap (channel-major: the
    // ConvTranspose2d parameter's own [Cout][k][k] order, so its weight gradient is a plain GEMM output)
    const int Wi = W * k, Hi = H * k;
    const int xx = p % Wi, yy = (p / Wi) % Hi, n = p / (Wi * Hi);
    const int ow = xx / k, kx = xx % k, oh = yy / k, ky = yy % k;
    const uint4 v = *reinterpret_cast<const uint4*>(x + (long long)p * ldx + c0);
    bf16* dst = y + (((long long)n * H + oh) * W + ow) * ldy;
    if (!chan_major) {
      *reinterpret_cast<uint4*>(dst + (ky * k + kx) * C + c0) = v;
    } else {
      const bf16* e = reinterpret_cast<const bf16*>(&v);
      const int kk = k * k, tap = ky * k + kx;
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[(long long)(c0 + j) * kk + tap] = e[j];
    }
  }
};

// 3x3 / pad 1 im2col of a few-channel image (C*9 <= 32): out[p][ci*9 + r*3 + s] = x[p + (r-1, s-1)][ci],
// zero outside the image and in the pad lanes.  Turns the K=27 first convolution of the U-Net into a 1x1
// convolution with one 64-wide K block instead of nine (one per tap, 61 of 64 lanes zero).
__global__ void im2col3x3_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ out, int N, int H, int W, int C) {
  const long long P = (long long)N * H * W;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int y2 = yy + r - 1, x2 = xx + t - 1;
        if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
          float f[8];
          load8(x + (p + (long long)(r - 1) * W + (t - 1)) * ldx, f);
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
            if (ci < C) v[ci * 9 + r * 3 + t] = f[ci];
        }
      }
    bf16* o = out + p * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float t8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) t8[i] = v[j + i];
      store8(o + j, t8);
    }
  }
}

__global__ void nchw_to_nhwc_kernel(const float* x, bf16* y, int ldy, int N, int C, int HW) {
  const long long P = (long long)N * HW;
  const int Cv = ldy / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * Cv; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % P; const int cv = (int)(i / P);
    const long long n = p / HW, hw = p % HW;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int c = cv * 8 + j; v[j] = c < C ? x[(n * C + c) * HW + hw] : 0.f; }
    store8(y + p * ldy + cv * 8, v);
  }
}
__global__ void nchw_to_nhwc_f32_kernel(const float* x, float* y, int ldy, int N, int C, int HW) {
  const long long P = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * ldy; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % P; const int c = (int)(i / P);
    const long long n = p / HW, hw = p % HW;
    y[p * ldy + c] = c < C ? x[(n * C + c) * HW + hw] : 0.f;
  }
}
// fp32 NHWC (pitch ldx) -> bf16 NHWC (pitch ldy), pad lanes zero
__global__ void cast_nhwc_kernel(const float* x, int ldx, bf16* y, int ldy, long long P, int C) {
  const int Cv = ldy / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * Cv; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / Cv; const int cv = (int)(i % Cv);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { const int c = cv * 8 + j; v[j] = c < C ? x[p * ldx + c] : 0.f; }
    store8(y + p * ldy + cv * 8, v);
  }
}
__global__ void nhwc_to_nchw_kernel(const void* x, int x_f32, int ldx, float* y, int N, int C, int HW) {
  const long long P = (long long)N * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * C; i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW; const int c = (int)((i / HW) % C); const long long n = i / ((long long)HW * C);
    const long long p = n * HW + hw;
    y[i] = x_f32 ? reinterpret_cast<const float*>(x)[p * ldx + c] : __bfloat162float(reinterpret_cast<const bf16*>(x)[p * ldx + c]);
  }
}

}  // namespace

// =========================================================================================== C ABI
#define ST ((cudaStream_t)stream)
#define REQ8(v, name) if ((v) & 7) { dm_set_error(name ": channel pitch must be a multiple of 8"); return DM_ERR_ARG; }

extern "C" int dm_nchw_to_nhwc(const float* x, void* y, int ldy, int y_f32, int N, int C, int H, int W, void* stream) {
  if (y_f32) {
    long long tot = (long long)N * H * W * ldy;
    nchw_to_nhwc_f32_kernel<<<ew_grid(tot), kEwThreads, 0, ST>>>(x, (float*)y, ldy, N, C, H * W);
  } else {
    REQ8(ldy, "dm_nchw_to_nhwc");
    long long tot = (long long)N * H * W * (ldy / 8);
    nchw_to_nhwc_kernel<<<ew_grid(tot), kEwThreads, 0, ST>>>(x, (bf16*)y, ldy, N, C, H * W);
  }
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_cast_nhwc(const float* x, int ldx, void* y, int ldy, long long P, int C, void* stream) {
  REQ8(ldy, "dm_cast_nhwc");
  cast_nhwc_kernel<<<ew_grid(P * (ldy / 8)), kEwThreads, 0, ST>>>(x, ldx, (bf16*)y, ldy, P, C);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_nhwc_to_nchw(const void* x, int x_f32, int ldx, float* y, int N, int C, int H, int W, void* stream) {
  long long tot = (long long)N * H * W * C;
  nhwc_to_nchw_kernel<<<ew_grid(tot), kEwThreads, 0, ST>>>(x, x_f32, ldx, y, N, C, H * W);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_im2col3x3(const void* x, int ldx, void* out, int N, int H, int W, int C, void* stream) {
  REQ8(ldx, "dm_im2col3x3");
  if (C < 1 || C > 3) { dm_set_error("dm_im2col3x3: 1..3 channels"); return DM_ERR_ARG; }
  im2col3x3_kernel<<<ew_grid((long long)N * H * W), kEwThreads, 0, ST>>>((const bf16*)x, ldx, (bf16*)out, N, H, W, C);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_space_to_depth(const void* x, int ldx, void* y, int ldy, int N, int H, int W, int C, int k, int chan_major,
                                 void* stream) {
  REQ8(ldx, "dm_space_to_depth"); REQ8(ldy, "dm_space_to_depth"); REQ8(C, "dm_space_to_depth");
  S2D f{(const bf16*)x, ldx, (bf16*)y, ldy, H, W, C, k, chan_major};
  return ew_launch((long long)N * H * k * W * k, C, f, ST);
}

extern "C" int dm_bn_finalize(const float* partials, int m_tiles, int ld, int C, double count, float* mean, float* invstd,
                              float* running_mean, float* running_var, float momentum, float eps, const float* conv_bias,
                              void* stream) {
  bn_finalize_kernel<<<dm::cdiv(C, 16), 1024, 0, ST>>>(partials, m_tiles, ld, C, count, mean, invstd, running_mean,
                                                       running_var, momentum, eps, conv_bias);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
static int bn_stats_blocks(long long P, int C) {
  const ChanMap m = chan_map(C);
  int gx = chan_grid_x(P, m, 16);
  const int cap = DM_NUM_SMS * 4 / m.cvt;
  return gx > cap ? (cap < 1 ? 1 : cap) : gx;
}
extern "C" int dm_bn_stats_rows(long long P, int C) { return bn_stats_blocks(P, C); }
extern "C" int dm_bn_stats(const void* y, int ldy, float* part, int ld, long long P, int C, void* stream) {
  REQ8(ldy, "dm_bn_stats");
  if (P <= 0 || P >= (1ll << 31)) { dm_set_error("dm_bn_stats: bad pixel count"); return DM_ERR_ARG; }
  const ChanMap m = chan_map(C);
  dim3 grid(bn_stats_blocks(P, C), m.cvt);
  bn_stats_kernel<<<grid, m.threads, (size_t)m.threads * 16 * sizeof(float), ST>>>((const bf16*)y, ldy, part, ld, (unsigned)P, C, m.VPB, m.R);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_bn_act_fwd(const void* y, int ldy, const float* mean, const float* invstd, const float* gamma,
                             const float* beta, void* z, int ldz, long long P, int C, int act, void* stream) {
  REQ8(ldy, "dm_bn_act_fwd"); REQ8(ldz, "dm_bn_act_fwd");
  if (P <= 0) return DM_OK;
  if (P >= (1ll << 31)) { dm_set_error("dm_bn_act_fwd: too many pixels"); return DM_ERR_ARG; }
  const ChanMap m = chan_map(C);
  dim3 grid(chan_grid_x(P, m, 8), m.cvt);
#define BN_FWD(A) bn_fwd_kernel<A><<<grid, m.threads, 0, ST>>>((const bf16*)y, ldy, mean, invstd, gamma, beta, (bf16*)z, ldz, (unsigned)P, C, m.VPB, m.R, 0)
  if (act == 1) BN_FWD(1); else if (act == 2) BN_FWD(2); else BN_FWD(0);
#undef BN_FWD
  DM_CHECK_LAUNCH();
  return DM_OK;
}
static int bn_bwd_blocks(long long P, int C) {
  // about one resident wave: few partial rows to finalize
  const ChanMap m = chan_map(C);
  int b = chan_grid_x(P, m, 16);
  const int cap = DM_NUM_SMS * 4 / m.cvt;
  return b > cap ? (cap < 1 ? 1 : cap) : b;
}
extern "C" long long dm_bn_act_bwd_scratch(long long P, int C) {
  return (long long)bn_bwd_blocks(P, C) * 2 * C + 3LL * C;
}
extern "C" int dm_bn_act_bwd(const void* dz, int lddz, const void* y, int ldy, const float* mean, const float* invstd,
                             const float* gamma, const float* beta, void* dy, int lddy, float* dgamma, float* dbeta,
                             float* dbias, float* scratch, long long P, int C, int act, int training, void* stream) {
  REQ8(lddz, "dm_bn_act_bwd"); REQ8(ldy, "dm_bn_act_bwd"); REQ8(lddy, "dm_bn_act_bwd");
  if (P <= 0) return DM_OK;
  if (P >= (1ll << 31)) { dm_set_error("dm_bn_act_bwd: too many pixels"); return DM_ERR_ARG; }
  const ChanMap m = chan_map(C);
  const int nblk = bn_bwd_blocks(P, C);
  float* coef = scratch;                    // [3][C]
  float* part = scratch + 3LL * C;          // [nblk][2][C]
  dim3 grid(nblk, m.cvt);
  const size_t smem = (size_t)m.threads * 16 * kRingS * kRingU * 2;       // cp.async ring (>= 64 B per thread for the block reduction)
#define BN_RED(A) bn_bwd_reduce_kernel<A><<<grid, m.threads, smem, ST>>>((const bf16*)dz, lddz, (const bf16*)y, ldy, mean, invstd, gamma, beta, part, (unsigned)P, C, m.VPB, m.R, 0)
  if (act == 1) BN_RED(1); else if (act == 2) BN_RED(2); else BN_RED(0);
#undef BN_RED
  DM_CHECK_LAUNCH();
  bn_bwd_finalize_kernel<<<dm::cdiv(C, 16), 1024, 0, ST>>>(part, nblk, C, (double)P, mean, invstd, gamma, coef, dgamma, dbeta,
                                                         dbias, training);
  DM_CHECK_LAUNCH();
  dim3 grid2(chan_grid_x(P, m, 4), m.cvt);
#define BN_APP(A) bn_bwd_apply_kernel<A><<<grid2, m.threads, smem, ST>>>((const bf16*)dz, lddz, (const bf16*)y, ldy, mean, invstd, gamma, beta, coef, (bf16*)dy, lddy, (unsigned)P, C, m.VPB, m.R, 0)
  if (act == 1) BN_APP(1); else if (act == 2) BN_APP(2); else BN_APP(0);
#undef BN_APP
  DM_CHECK_LAUNCH();
  return DM_OK;
}

static int gn_blocks(int N, long long HW, int C, int vec) {
  // blocks per sample: enough CTAs in total for a full wave, but few partial rows to fold
  const ChanMap m = chan_map(C, vec);
  int b = chan_grid_x(HW, m, 16);
  int cap = DM_NUM_SMS * 4 / (m.cvt * (N > 0 ? N : 1)); if (cap < 1) cap = 1;
  return b > cap ? cap : b;
}
/* scratch layout (floats): sc[N*C] sh[N*C] sx[N*C] coef[3*N*C] part[N*blocks*2*C] */
extern "C" long long dm_gn_scratch(int N, int HW, int C) {
  const int b8 = gn_blocks(N, HW, C, 8), b4 = gn_blocks(N, HW, C, 4);
  return 6LL * N * C + 2LL * N * (b8 > b4 ? b8 : b4) * C;
}
extern "C" int dm_gn_act_fwd(const void* x, int ldx, const float* gamma, const float* beta, void* z, int ldz, float* mean,
                             float* rstd, float* scratch, int N, int HW, int C, int G, float eps, int act, void* stream) {
  REQ8(ldx, "dm_gn_act_fwd"); REQ8(ldz, "dm_gn_act_fwd");
  if (C % G || C / G > 512) { dm_set_error("dm_gn_act_fwd: C must be divisible by G with at most 512 channels per group"); return DM_ERR_ARG; }
  if (N <= 0 || HW <= 0) return DM_OK;
  if (N > 65535) { dm_set_error("dm_gn_act_fwd: batch too large"); return DM_ERR_ARG; }
  float *sc = scratch, *sh = sc + (long long)N * C, *sx = sh + (long long)N * C, *part = scratch + 6LL * N * C;
  const ChanMap m = chan_map(C);
  const int nblk = gn_blocks(N, HW, C, 8);
  bn_stats_kernel<<<dim3(nblk, m.cvt, N), m.threads, (size_t)m.threads * 16 * sizeof(float), ST>>>((const bf16*)x, ldx, part, C,
                                                                                                  (unsigned)HW, C, m.VPB, m.R);
  DM_CHECK_LAUNCH();
  gn_fold_fwd_kernel<<<N * G, 256, 0, ST>>>(part, nblk, C, G, (double)HW * (C / G), eps, gamma, beta, mean, rstd, sc, sh, sx);
  DM_CHECK_LAUNCH();
  dim3 grid(chan_grid_x(HW, m, 8), m.cvt, N);
  { int cap = DM_NUM_SMS * 8 / (m.cvt * N); if (cap < 1) cap = 1; if ((int)grid.x > cap) grid.x = cap; }
  // z = act(x * sc[n][c] + sh[n][c]): the BatchNorm apply kernel on per-sample coefficient rows (rows = 1: invstd -> sc, mean -> sh)
#define GN_FWD(A) bn_fwd_kernel<A><<<grid, m.threads, 0, ST>>>((const bf16*)x, ldx, sh, sc, nullptr, nullptr, (bf16*)z, ldz, (unsigned)HW, C, m.VPB, m.R, 1)
  if (act == 1) GN_FWD(1); else if (act == 2) GN_FWD(2); else GN_FWD(0);
#undef GN_FWD
  DM_CHECK_LAUNCH();
  return DM_OK;
}
/* scratch: the forward call's buffer (sc/sh rows are read, coef/part are overwritten) */
extern "C" int dm_gn_act_bwd(const void* dz, int lddz, const void* x, int ldx, const float* mean, const float* rstd,
                             const float* gamma, const float* beta, void* dx, int lddx, float* dgamma, float* dbeta,
                             float* scratch, int N, int HW, int C, int G, int act, void* stream) {
  REQ8(lddz, "dm_gn_act_bwd"); REQ8(ldx, "dm_gn_act_bwd"); REQ8(lddx, "dm_gn_act_bwd");
  if (N <= 0 || HW <= 0) return DM_OK;
  float *sc = scratch, *sh = sc + (long long)N * C, *coef = scratch + 3LL * N * C, *part = scratch + 6LL * N * C;
  const ChanMap m = chan_map(C);
  const int nblk = gn_blocks(N, HW, C, 8);
  const size_t smem = (size_t)m.threads * 16 * kRingS * kRingU * 2;       // cp.async ring of the two-operand kernels
#define GN_RED(A) bn_bwd_reduce_kernel<A><<<dim3(nblk, m.cvt, N), m.threads, smem, ST>>>((const bf16*)dz, lddz, (const bf16*)x, ldx, sh, sc, nullptr, nullptr, part, (unsigned)HW, C, m.VPB, m.R, 1)
  if (act == 1) GN_RED(1); else if (act == 2) GN_RED(2); else GN_RED(0);
#undef GN_RED
  DM_CHECK_LAUNCH();
  gn_fold_bwd_kernel<<<N * G, 256, 0, ST>>>(part, nblk, C, G, (double)HW * (C / G), gamma, mean, rstd, coef, dgamma, dbeta);
  DM_CHECK_LAUNCH();
  dim3 grid(chan_grid_x(HW, m, 4), m.cvt, N);
  { int cap = DM_NUM_SMS * 8 / (m.cvt * N); if (cap < 1) cap = 1; if ((int)grid.x > cap) grid.x = cap; }
#define GN_APP(A) bn_bwd_apply_kernel<A><<<grid, m.threads, smem, ST>>>((const bf16*)dz, lddz, (const bf16*)x, ldx, sh, sc, nullptr, nullptr, coef, (bf16*)dx, lddx, (unsigned)HW, C, m.VPB, m.R, 1)
  if (act == 1) GN_APP(1); else if (act == 2) GN_APP(2); else GN_APP(0);
#undef GN_APP
  DM_CHECK_LAUNCH();
  (void)beta;
  return DM_OK;
}

extern "C" int dm_pool_nhw(const void* x, int ldx, float* out, int N, int HW, int C, float scale, void* stream) {
  REQ8(ldx, "dm_pool_nhw");
  RedArgs A{};
  A.a = (const bf16*)x; A.a_hi = (long long)HW * ldx; A.a_ps = ldx; A.gdiv = 1; A.count = HW; A.C = C; A.mode = 0;
  A.scale = scale; A.out1 = out; A.G = 1; A.stat_div = 1;
  return launch_reduce(A, N, ST);
}
extern "C" int dm_pool_prod_nhw(const void* a, int lda, const void* b, int ldb, float* out, int N, int HW, int C,
                                float scale, void* stream) {
  REQ8(lda, "dm_pool_prod_nhw"); REQ8(ldb, "dm_pool_prod_nhw");
  RedArgs A{};
  A.a = (const bf16*)a; A.b = (const bf16*)b;
  A.a_hi = (long long)HW * lda; A.a_ps = lda; A.b_hi = (long long)HW * ldb; A.b_ps = ldb;
  A.gdiv = 1; A.count = HW; A.C = C; A.mode = 2; A.scale = scale; A.out1 = out; A.out2 = nullptr; A.G = 1; A.stat_div = 1;
  return launch_reduce(A, N, ST);
}
extern "C" int dm_colsum(const void* dy, int lddy, float* db, long long P, int C, void* stream) {
  REQ8(lddy, "dm_colsum");
  if (P <= 0) return DM_OK;
  if (P >= (1ll << 31)) { dm_set_error("dm_colsum: too many pixels"); return DM_ERR_ARG; }
  const ChanMap m = chan_map(C);
  int gx = chan_grid_x(P, m, 16);
  int cap = DM_NUM_SMS * 5 / m.cvt; if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  if ((long long)gx * 2 * C > g_ws_floats) {
    if (g_ws_floats < 2LL * C) { dm_set_error("dm_colsum needs scratch: call dm_set_workspace() first"); return DM_ERR_ARG; }
    gx = (int)(g_ws_floats / (2LL * C));
  }
  dim3 grid(gx, m.cvt);
  colsum_kernel<<<grid, m.threads, (size_t)m.threads * 8 * sizeof(float), ST>>>((const bf16*)dy, lddy, g_ws, (unsigned)P, C, m.VPB, m.R);
  DM_CHECK_LAUNCH();
  reduce_finalize_kernel<<<finalize_grid(C), 1024, 0, ST>>>(g_ws, gx, 1, C, 1.0f, db, nullptr, 1);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_set_workspace(void* ptr, long long bytes) {
  g_ws = reinterpret_cast<float*>(ptr);
  g_ws_floats = ptr ? bytes / 4 : 0;
  return DM_OK;
}
extern "C" int dm_se_apply_fwd(const void* x2, int ld2, const float* gate, const void* res, int ldr, void* out, int ldo,
                               int N, int HW, int C, float scale, void* stream) {
  REQ8(ld2, "dm_se_apply_fwd"); REQ8(ldr, "dm_se_apply_fwd"); REQ8(ldo, "dm_se_apply_fwd");
  SeFwd f{(const bf16*)x2, ld2, gate, (const bf16*)res, ldr, (bf16*)out, ldo, C, HW, scale};
  return ew_launch((long long)N * HW, C, f, ST);
}
extern "C" int dm_se_apply_bwd(const void* dout, int lddo, const float* gate, const float* dpool, void* dx2, int lddx2,
                               void* dres, int lddr, int N, int HW, int C, float scale, void* stream) {
  REQ8(lddo, "dm_se_apply_bwd"); REQ8(lddx2, "dm_se_apply_bwd"); REQ8(lddr, "dm_se_apply_bwd");
  SeBwd f{(const bf16*)dout, lddo, gate, dpool, (bf16*)dx2, lddx2, (bf16*)dres, lddr, C, HW, scale, 1.0f / (float)HW};
  return ew_launch((long long)N * HW, C, f, ST);
}

extern "C" int dm_ca_pool(const void* a, int lda, const void* b, int ldb, float* oh, float* ow, int N, int H, int W, int C,
                          float scale_h, float scale_w, void* stream) {
  REQ8(lda, "dm_ca_pool"); if (b) REQ8(ldb, "dm_ca_pool");
  int rc;
  RedArgs A{};
  A.a = (const bf16*)a; A.b = (const bf16*)b; A.C = C; A.mode = b ? 2 : 0; A.G = 1; A.stat_div = 1;
  // rows: group g = n*H + h, pixels along w
  A.gdiv = 1; A.a_hi = (long long)W * lda; A.a_lo = 0; A.a_ps = lda; A.b_hi = (long long)W * ldb; A.b_lo = 0; A.b_ps = ldb;
  A.count = W; A.scale = scale_h; A.out1 = oh; A.out2 = nullptr;
  rc = launch_reduce(A, N * H, ST); if (rc) return rc;
  // columns: group g = n*W + w, pixels along h (stride W)
  A.gdiv = W; A.a_hi = (long long)H * W * lda; A.a_lo = lda; A.a_ps = (long long)W * lda;
  A.b_hi = (long long)H * W * ldb; A.b_lo = ldb; A.b_ps = (long long)W * ldb;
  A.count = H; A.scale = scale_w; A.out1 = ow;
  return launch_reduce(A, N * W, ST);
}
extern "C" int dm_ca_gate_fwd(const void* x, int ldx, const float* ah, const float* aw, void* out, int ldo, int N, int H,
                              int W, int C, void* stream) {
  REQ8(ldx, "dm_ca_gate_fwd"); REQ8(ldo, "dm_ca_gate_fwd");
  CaFwd f{(const bf16*)x, ldx, ah, aw, (bf16*)out, ldo, C, H, W};
  return ew_launch((long long)N * H * W, C, f, ST);
}
extern "C" int dm_ca_gate_bwd(const void* dout, int lddo, const float* ah, const float* aw, const float* dxh,
                              const float* dxw, void* dx, int lddx, int N, int H, int W, int C, void* stream) {
  REQ8(lddo, "dm_ca_gate_bwd"); REQ8(lddx, "dm_ca_gate_bwd");
  CaBwd f{(const bf16*)dout, lddo, ah, aw, dxh, dxw, (bf16*)dx, lddx, C, H, W, 1.0f / (float)W, 1.0f / (float)H};
  return ew_launch((long long)N * H * W, C, f, ST);
}

extern "C" int dm_upcat_fwd_shared(const void* a, int lda, int Ca, const void* b, int ldb, int Cb, int Nb, void* out, int ldo,
                                   int N, int h, int w, void* stream);
extern "C" int dm_upcat_fwd(const void* a, int lda, int Ca, const void* b, int ldb, int Cb, void* out, int ldo, int N,
                            int h, int w, void* stream) {
  return dm_upcat_fwd_shared(a, lda, Ca, b, ldb, Cb, 0, out, ldo, N, h, w, stream);
}
/* b holds Nb samples that the N = k * Nb samples of a share cyclically (sample n reads b[n % Nb]); Nb = 0: b holds N */
extern "C" int dm_upcat_fwd_shared(const void* a, int lda, int Ca, const void* b, int ldb, int Cb, int Nb, void* out, int ldo,
                                   int N, int h, int w, void* stream) {
  REQ8(lda, "dm_upcat_fwd"); REQ8(ldb, "dm_upcat_fwd"); REQ8(ldo, "dm_upcat_fwd"); REQ8(Ca, "dm_upcat_fwd(Ca)");
  if (Nb < 0 || (Nb > 0 && N % Nb)) { dm_set_error("dm_upcat_fwd_shared: N must be a multiple of Nb"); return DM_ERR_ARG; }
  UpcatFwd f{(const bf16*)a, lda, Ca, (const bf16*)b, ldb, Cb, (bf16*)out, ldo, h, w,
             h > 1 ? (float)(h - 1) / (float)(2 * h - 1) : 0.f, w > 1 ? (float)(w - 1) / (float)(2 * w - 1) : 0.f, Nb == N ? 0 : Nb};
  if (dm_debug_value(9) == 1) return ew_launch((long long)N * 4 * h * w, Ca + Cb, f, ST);   // dev: one thread per output pixel
  const long long Cv = (Ca + Cb + 7) / 8, total = (long long)N * (h + 1) * (w + 1) * Cv;
  if (total <= 0) return DM_OK;
  if (total >= (1ll << 32) || (long long)N * 4 * h * w * Cv >= (1ll << 32)) { dm_set_error("dm_upcat_fwd: tensor too large"); return DM_ERR_ARG; }
  if ((long long)N * (h + 1) > 65535) { dm_set_error("dm_upcat_fwd: N * (h + 1) must fit the grid's y extent (65535)"); return DM_ERR_ARG; }
  const dim3 grid((unsigned)(((w + 1) * Cv + kEwThreads - 1) / kEwThreads), (unsigned)(N * (h + 1)));
  upcat_fwd_quad_kernel<<<grid, kEwThreads, 0, ST>>>(f, (unsigned)Cv);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_upcat_bwd(const void* dout, int lddo, void* da, int ldda, int Ca, void* db, int lddb, int Cb, int N,
                            int h, int w, void* stream) {
  REQ8(lddo, "dm_upcat_bwd"); REQ8(ldda, "dm_upcat_bwd"); REQ8(lddb, "dm_upcat_bwd"); REQ8(Ca, "dm_upcat_bwd(Ca)");
  UpcatBwd f{(const bf16*)dout, lddo, (bf16*)da, ldda, Ca, (bf16*)db, lddb, Cb, h, w,
             h > 1 ? (float)(h - 1) / (float)(2 * h - 1) : 0.f, w > 1 ? (float)(w - 1) / (float)(2 * w - 1) : 0.f};
  if (dm_debug_value(9) == 1) return ew_launch((long long)N * h * w, Ca + Cb, f, ST);      // dev: one thread per input pixel
  const long long Cv = (Ca + Cb + 7) / 8, total = (long long)N * ((h + 1) / 2) * ((w + 1) / 2) * Cv;
  if (total <= 0) return DM_OK;
  if ((long long)N * 4 * h * w * Cv >= (1ll << 32)) { dm_set_error("dm_upcat_bwd: tensor too large"); return DM_ERR_ARG; }
  // one work item per thread in 128-thread blocks: at ~100 registers that is 5 resident blocks (20 warps) per SM instead
  // of 2 x 256 threads, and no capped grid-stride loop (786 K items on 606 K threads ran as one full round + a 30 % one)
  upcat_bwd_quad_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ST>>>(f, (unsigned)total, (unsigned)Cv);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_film_fwd(const void* x, int ldx, const float* ce, const float* te, void* out, int ldo, int N, int HW,
                           int C, void* stream) {
  REQ8(ldx, "dm_film_fwd"); REQ8(ldo, "dm_film_fwd");
  FilmFwd f{(const bf16*)x, ldx, ce, te, (bf16*)out, ldo, C, HW};
  return ew_launch((long long)N * HW, C, f, ST);
}
extern "C" int dm_film_bwd(const void* dout, int lddo, const void* x, int ldx, const float* ce, void* dx, int lddx,
                           float* dce, float* dte, int N, int HW, int C, void* stream) {
  REQ8(lddo, "dm_film_bwd"); REQ8(ldx, "dm_film_bwd"); REQ8(lddx, "dm_film_bwd");
  int rc;
  RedArgs A{};
  A.a = (const bf16*)dout; A.b = (const bf16*)x;
  A.a_hi = (long long)HW * lddo; A.a_ps = lddo; A.b_hi = (long long)HW * ldx; A.b_ps = ldx;
  A.gdiv = 1; A.count = HW; A.C = C; A.mode = 2; A.scale = 1.f; A.out1 = dce; A.out2 = dte; A.G = 1; A.stat_div = 1;
  rc = launch_reduce(A, N, ST); if (rc) return rc;
  ScaleNC f{(const bf16*)dout, lddo, ce, (bf16*)dx, lddx, C, HW};
  return ew_launch((long long)N * HW, C, f, ST);
}

extern "C" int dm_avgpool_act_fwd(const void* x, int ldx, void* out, int ldo, int N, int H, int W, int C, int k, int act,
                                  void* stream) {
  REQ8(ldx, "dm_avgpool_act_fwd"); REQ8(ldo, "dm_avgpool_act_fwd");
  AvgPoolFwd f{(const bf16*)x, ldx, (bf16*)out, ldo, H, W, C, k, act};
  return ew_launch((long long)N * (H / k) * (W / k), C, f, ST);
}
extern "C" int dm_avgpool_act_bwd(const void* dout, int lddo, const void* x, int ldx, void* dx, int lddx, int N, int H,
                                  int W, int C, int k, int act, void* stream) {
  REQ8(lddo, "dm_avgpool_act_bwd"); REQ8(ldx, "dm_avgpool_act_bwd"); REQ8(lddx, "dm_avgpool_act_bwd");
  AvgPoolBwd f{(const bf16*)dout, lddo, (const bf16*)x, ldx, (bf16*)dx, lddx, H, W, C, k, act};
  return ew_launch((long long)N * (H / k) * (W / k), C, f, ST);
}
extern "C" int dm_maxpool2_fwd(const void* x, int ldx, void* out, int ldo, int N, int H, int W, int C, void* stream) {
  REQ8(ldx, "dm_maxpool2_fwd"); REQ8(ldo, "dm_maxpool2_fwd");
  MaxPoolFwd f{(const bf16*)x, ldx, (bf16*)out, ldo, H, W, C};
  return ew_launch((long long)N * (H / 2) * (W / 2), C, f, ST);
}
extern "C" int dm_maxpool2_bwd(const void* dout, int lddo, const void* x, int ldx, void* dx, int lddx, int N, int H, int W,
                               int C, void* stream) {
  REQ8(lddo, "dm_maxpool2_bwd"); REQ8(ldx, "dm_maxpool2_bwd"); REQ8(lddx, "dm_maxpool2_bwd");
  MaxPoolBwd f{(const bf16*)dout, lddo, (const bf16*)x, ldx, (bf16*)dx, lddx, H, W, C};
  return ew_launch((long long)N * (H / 2) * (W / 2), C, f, ST);
}

extern "C" int dm_mask_fma(const void* x, int ldx, const void* y, int ldy, const float* mask, float thresh, void* out,
                           int ldo, long long P, int C, void* stream) {
  if (x) REQ8(ldx, "dm_mask_fma"); REQ8(ldy, "dm_mask_fma"); REQ8(ldo, "dm_mask_fma");
  MaskFma f{(const bf16*)x, ldx, (const bf16*)y, ldy, mask, thresh, (bf16*)out, ldo, C};
  return ew_launch(P, C, f, ST);
}
extern "C" int dm_axpby(const void* a, int lda, const void* b, int ldb, void* out, int ldo, long long P, int C, float sa,
                        float sb, void* stream) {
  REQ8(lda, "dm_axpby"); if (b) REQ8(ldb, "dm_axpby"); REQ8(ldo, "dm_axpby");
  Axpby f{(const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)out, ldo, C, sa, sb};
  return ew_launch(P, C, f, ST);
}
