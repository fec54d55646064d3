// CoordAttn gate network (new_scripy.py:97-140) as two forward and six backward kernels instead of ~90 tiny
// library launches per instance and direction pair.  All fp32; R = N*L rows per direction (L = H = W), C channels,
// m = C/16 hidden channels:
//   U_d  = X_d W1_d^T + b1_d                               d in {h, w},  X_d = directional means [R][C]
//   T0_d = gelu(BN_d(U_d))                                  BatchNorm over the R rows (batch or running statistics)
//   T_h  = T0_h + sigmoid(gamma_h) * (T0_w Wwh^T + bwh)     cross interaction, same row index (H == W)
//   T_w  = T0_w + sigmoid(gamma_w) * (T0_h Whw^T + bhw)
//   A_d  = k_d * sigmoid(T_d Wc_d^T + bc_d)                 k_h = a/(a+b+1e-8), k_w = b/(a+b+1e-8), a = sigmoid(alpha) ...
// The C x m contractions are CUDA-core dot products (0.005 % of the model's FLOPs): a block owns kRows rows, keeps its
// X / dZ tile in shared memory and streams the weight matrices from L2.  Cross-block reductions: BatchNorm sums go
// through per-block partial rows (fixed order); the two C x m weight gradients come from a kernel that is parallel over
// outputs and loops over the rows, only the m x m projection gradients use fp32 atomics.
#include <stdio.h>

#include "common.cuh"
#include "dm_b200.h"

namespace {

constexpr int kRows = 8;          // rows per block
constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_exact(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// out1[j] / out2[j] = sum over the nblk partial rows of the two statistics of direction d (row layout [d][2][m]).
// Threads split the rows into slices, so no thread walks a long chain of dependent L2 round trips; fixed order.
// scratch: 2 * kThreads floats; m <= kThreads.  Ends with a __syncthreads().
__device__ __forceinline__ void sum_partials(const float* __restrict__ part, int nblk, int m, int d, float* scratch,
                                             float* out1, float* out2) {
  const int slices = kThreads / m;
  const int j = threadIdx.x % m, slice = threadIdx.x / m;
  float s1 = 0.0f, s2 = 0.0f;
  if (slice < slices)
    for (int b = slice; b < nblk; b += slices) {
      const float* row = part + (((long long)b * 2 + d) * 2) * m;
      s1 += __ldg(row + j); s2 += __ldg(row + m + j);
    }
  __syncthreads();                                   // scratch may still be read by a previous call
  scratch[threadIdx.x] = s1; scratch[kThreads + threadIdx.x] = s2;
  __syncthreads();
  for (int jj = threadIdx.x; jj < m; jj += kThreads) {
    double a = 0.0, b = 0.0;
    for (int sl = 0; sl < slices; ++sl) { a += (double)scratch[sl * m + jj]; b += (double)scratch[kThreads + sl * m + jj]; }
    out1[jj] = (float)a; out2[jj] = (float)b;
  }
  __syncthreads();
}

struct Scalars { float sg[2], k[2], a, b, S; };
__device__ __forceinline__ Scalars load_scalars(const DmCaGates& p) {
  Scalars s;
  s.sg[0] = sigmoidf_(p.gamma_h[0]); s.sg[1] = sigmoidf_(p.gamma_w[0]);
  s.a = sigmoidf_(p.alpha[0]); s.b = sigmoidf_(p.beta[0]);
  s.S = s.a + s.b + 1e-8f;
  s.k[0] = s.a / s.S; s.k[1] = s.b / s.S;
  return s;
}

// ---- forward 1: U_d = X_d W1_d^T + b1_d and the per-block BatchNorm partial sums.  grid (nblk, 2, ceil(m/8)): every
// block streams only its own eight rows of W1 (a single SM cannot pull a whole weight matrix fast enough)
__global__ void __launch_bounds__(kThreads) ca_lin1_kernel(const DmCaGates p) {
  extern __shared__ float xs[];          // [kRows][C]
  const int d = blockIdx.y, r0 = blockIdx.x * kRows, C = p.C, m = p.m, R = p.R;
  const float* X = d ? p.xw : p.xh;
  const float* W1 = d ? p.w1_w : p.w1_h;
  const float* B1 = d ? p.b1_w : p.b1_h;
  for (int i = threadIdx.x; i < kRows * C; i += kThreads) {
    const int r = i / C, c = i - r * C;
    xs[i] = (r0 + r < R) ? X[(long long)(r0 + r) * C + c] : 0.0f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {                                        // one hidden channel per warp; blockIdx.z selects the group of eight
    const int j = blockIdx.z * (kThreads / 32) + warp;
    if (j >= m) return;
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
    const float* w = W1 + (long long)j * C;
#pragma unroll 4
    for (int c = lane; c < C; c += 32) {
      const float wv = __ldg(w + c);
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = fmaf(xs[r * C + c], wv, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = dm::warp_sum(acc[r]);
    if (lane == 0) {
      const float b = B1[j];
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (r0 + r < R) {
          const float u = acc[r] + b;
          p.u[((long long)d * R + r0 + r) * m + j] = u;
          s1 += u; s2 = fmaf(u, u, s2);
        }
      float* part = p.part + (((long long)blockIdx.x * 2 + d) * 2) * m;
      part[j] = s1; part[m + j] = s2;
    }
  }
}

// Shared by forward 2 and backward 1: statistics -> shared memory, then Hhat, T0 = gelu(Hhat), P_d = T0_d Wp_d^T + bp_d.
// sm layout (floats): mean[2][m] rstd[2][m] hh[2][kRows][m] t0[2][kRows][m] pp[2][kRows][m]
struct GateSmem { float *mean, *rstd, *hh, *t0, *pp; };
__device__ __forceinline__ GateSmem carve_gate(float* sm, int m) {
  GateSmem g;
  g.mean = sm; g.rstd = sm + 2 * m; g.hh = sm + 4 * m; g.t0 = g.hh + 2 * kRows * m; g.pp = g.t0 + 2 * kRows * m;
  return g;
}
__device__ __forceinline__ void gate_front(const DmCaGates& p, const GateSmem& g, int r0, bool write_stats) {
  const int m = p.m, R = p.R;
  __shared__ float scratch[2 * kThreads];
  float* sums = g.hh;                       // [2][2m] scratch until Hhat is written below (needs 4m <= 2*kRows*m)
  if (write_stats && p.training) {
    sum_partials(p.part, p.nblk, m, 0, scratch, sums, sums + 2 * m);
    sum_partials(p.part, p.nblk, m, 1, scratch, sums + m, sums + 3 * m);
  }
  for (int idx = threadIdx.x; idx < 2 * m; idx += kThreads) {
    const int d = idx / m, j = idx - d * m;
    float* rm = d ? p.bn_rm_w : p.bn_rm_h;
    float* rv = d ? p.bn_rv_w : p.bn_rv_h;
    float mean, rstd;
    if (write_stats && p.training) {
      const double s1 = (double)sums[idx], s2 = (double)sums[2 * m + idx];
      const double mu = s1 / R;
      double var = s2 / R - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = (float)mu; rstd = (float)(1.0 / sqrt(var + (double)p.eps));
      if (blockIdx.x == 0 && blockIdx.y == 0) {
        const double unb = R > 1 ? var * R / (R - 1.0) : var;
        rm[j] = (1.0f - p.momentum) * rm[j] + p.momentum * mean;
        rv[j] = (1.0f - p.momentum) * rv[j] + p.momentum * (float)unb;
      }
    } else if (write_stats) {
      mean = rm[j]; rstd = 1.0f / sqrtf(rv[j] + p.eps);
    } else {
      mean = p.stat[d * 2 * m + j]; rstd = p.stat[d * 2 * m + m + j];
    }
    if (write_stats && blockIdx.x == 0 && blockIdx.y == 0) { p.stat[d * 2 * m + j] = mean; p.stat[d * 2 * m + m + j] = rstd; }
    g.mean[idx] = mean; g.rstd[idx] = rstd;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * kRows * m; idx += kThreads) {
    const int d = idx / (kRows * m), rj = idx - d * kRows * m, r = rj / m, j = rj - r * m;
    float hh = 0.0f, t0 = 0.0f;
    if (r0 + r < R) {
      const float u = p.u[((long long)d * R + r0 + r) * m + j];
      const float ga = (d ? p.bn_g_w : p.bn_g_h)[j], be = (d ? p.bn_b_w : p.bn_b_h)[j];
      hh = (u - g.mean[d * m + j]) * g.rstd[d * m + j] * ga + be;
      t0 = gelu_exact(hh);
    }
    g.hh[idx] = hh; g.t0[idx] = t0;
    if (write_stats && blockIdx.y == 0 && r0 + r < R) p.t0[((long long)d * R + r0 + r) * m + j] = t0;
  }
  __syncthreads();
  // P_d[r][j] = bp_d[j] + sum_k T0_d[r][k] * wp_d[j][k]: a warp per (d, j), lanes over k (coalesced weight rows; one
  // lane per row of wp_d would turn every load into 32 separate sectors).  A warp takes kProjRows rows at a time and
  // requests all of their weights before the first use: one round trip to L2/DRAM per four rows (a rolled k loop
  // inside a rolled row loop paid m/32 dependent round trips per row -- 72 in a row for m = 96, most of the kernel).
  {
    constexpr int kProjRows = 4, kKU = kThreads / 32;         // m <= kThreads: at most kKU weights per lane and row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int dj0 = warp * kProjRows; dj0 < 2 * m; dj0 += (kThreads / 32) * kProjRows) {
      float wv[kProjRows][kKU];
#pragma unroll
      for (int q = 0; q < kProjRows; ++q) {
        const int dj = dj0 + q, d = dj >= m, j = dj - d * m;
        const float* wp = (d ? p.wp_w2h : p.wp_h2w) + (long long)j * m;
#pragma unroll
        for (int u = 0; u < kKU; ++u) {
          const int k = lane + 32 * u;
          wv[q][u] = (dj < 2 * m && k < m) ? __ldg(wp + k) : 0.0f;
        }
      }
#pragma unroll
      for (int q = 0; q < kProjRows; ++q) {
        const int dj = dj0 + q, d = dj >= m, j = dj - d * m;
        if (dj >= 2 * m) continue;                             // warp-uniform
        const float* t0 = g.t0 + d * kRows * m;
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
#pragma unroll
        for (int u = 0; u < kKU; ++u) {
          const int k = lane + 32 * u;
          if (k < m) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r] = fmaf(t0[r * m + k], wv[q][u], acc[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = dm::warp_sum(acc[r]);
        const float b = (d ? p.bp_w2h : p.bp_h2w)[j];
        float mine = 0.0f;
#pragma unroll
        for (int r = 0; r < kRows; ++r)
          if (lane == r) mine = acc[r];
        if (lane < kRows) g.pp[(d * kRows + lane) * m + j] = mine + b;      // pp[0] = h2w_proj(T0_h) feeds T_w, pp[1] feeds T_h
      }
    }
  }
  __syncthreads();
}

// ---- forward 2: statistics, gelu, cross interaction, output gates.  grid (nblk, ceil(C/256)): the small front part
// is recomputed per channel chunk, each block reads only its 256 rows of the output weights
__global__ void __launch_bounds__(kThreads) ca_gate_fwd_kernel(const DmCaGates p, const int use_slab) {
  extern __shared__ __align__(16) float sm[];
  const int m = p.m, C = p.C, R = p.R, r0 = blockIdx.x * kRows;
  const GateSmem g = carve_gate(sm, m);
  float* ts = g.pp + 2 * kRows * m;            // T[2][kRows][m]
  gate_front(p, g, r0, true);
  const Scalars s = load_scalars(p);
  for (int idx = threadIdx.x; idx < 2 * kRows * m; idx += kThreads) {
    const int d = idx / (kRows * m), rj = idx - d * kRows * m, r = rj / m;
    const float t = g.t0[idx] + s.sg[d] * g.pp[(1 - d) * kRows * m + rj];
    ts[idx] = t;
    if (r0 + r < R && blockIdx.y == 0) p.t[((long long)d * R + r0 + r) * m + (rj - r * m)] = t;
  }
  __syncthreads();
  for (int d = 0; d < 2; ++d) {
    const float* wc = d ? p.wc_w : p.wc_h;
    const float* bc = d ? p.bc_w : p.bc_h;
    float* out = d ? p.aw : p.ah;
    const float* t = ts + d * kRows * m;
    // a thread per output channel.  Its weight row (m contiguous floats) is first staged by the whole warp with
    // coalesced loads into a padded shared-memory slab: read straight from global, every load instruction of the
    // j loop would touch 32 different rows = 32 L1 wavefronts
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.y * kThreads + threadIdx.x;
    const float* w = wc + (long long)c * m;
    float* slab = ts + 2 * kRows * m + warp * 32 * (m + 1);
    if (use_slab) {
      const int cb = blockIdx.y * kThreads + warp * 32;
      for (int row = 0; row < 32 && cb + row < C; ++row)
        for (int j = lane; j < m; j += 32) slab[row * (m + 1) + j] = __ldg(wc + (long long)(cb + row) * m + j);
      __syncwarp();
    }
    if (c < C) {
      float acc[kRows];
      const float b = bc[c];
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = b;
      const float* ws = slab + lane * (m + 1);
      if (!use_slab && (m & 3) == 0 && (reinterpret_cast<uintptr_t>(wc) & 15) == 0) {
        // the thread's own weight row, eight 16-byte loads requested at a time (same j order as the scalar loop)
        const float4* w4 = reinterpret_cast<const float4*>(w);
        for (int j0 = 0; j0 < m; j0 += 32) {
          float4 wv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            wv[u] = j0 + 4 * u < m ? __ldg(w4 + (j0 >> 2) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = j0 + 4 * u;
            if (j >= m) continue;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
              const float4 tv = *reinterpret_cast<const float4*>(t + r * m + j);
              acc[r] = fmaf(tv.x, wv[u].x, acc[r]); acc[r] = fmaf(tv.y, wv[u].y, acc[r]);
              acc[r] = fmaf(tv.z, wv[u].z, acc[r]); acc[r] = fmaf(tv.w, wv[u].w, acc[r]);
            }
          }
        }
      } else {
#pragma unroll 4
        for (int j = 0; j < m; ++j) {
          const float wv = use_slab ? ws[j] : __ldg(w + j);
#pragma unroll
          for (int r = 0; r < kRows; ++r) acc[r] = fmaf(t[r * m + j], wv, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r)
        if (r0 + r < R) out[(long long)(r0 + r) * C + c] = s.k[d] * sigmoidf_(acc[r]);
    }
    __syncwarp();
  }
}

// ---- backward 1a: dZ = dA * k * a(1-a) (stored for the weight-gradient kernel), the scale sums, and this channel
// chunk's share of dT_d = dZ_d Wc_d (atomics into the zeroed q.dt).  grid (nblk, ceil(C/256), 2)
__global__ void __launch_bounds__(kThreads) ca_gate_bwd_a_kernel(const DmCaGates p, const DmCaGatesGrad q) {
  extern __shared__ __align__(16) float sm[];
  const int m = p.m, C = p.C, R = p.R, r0 = blockIdx.x * kRows, c0 = blockIdx.y * kThreads, d = blockIdx.z;
  float* dzs = sm;                       // [kRows][kThreads]
  float* dts = dzs + kRows * kThreads;   // [kRows][m]
  __shared__ float red[kThreads / 32];
  const Scalars s = load_scalars(p);
  const float* A = d ? p.aw : p.ah;
  const float* dA = d ? q.d_aw : q.d_ah;
  const float* wc = d ? p.wc_w : p.wc_h;
  const float kd = s.k[d], inv_k = 1.0f / kd;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRows * m; i += kThreads) dts[i] = 0.0f;
  const int c = c0 + threadIdx.x;
  float dk = 0.0f;
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    float dz = 0.0f;
    if (c < C && r0 + r < R) {
      const float a = A[(long long)(r0 + r) * C + c] * inv_k, go = dA[(long long)(r0 + r) * C + c];
      dk = fmaf(go, a, dk);
      dz = go * kd * a * (1.0f - a);
      q.dz[((long long)d * R + r0 + r) * C + c] = dz;
    }
    dzs[r * kThreads + threadIdx.x] = dz;
  }
  dk = dm::warp_sum(dk);
  if (lane == 0) red[warp] = dk;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.0f;
    for (int w = 0; w < kThreads / 32; ++w) tot += red[w];
    atomicAdd(q.scal + d, tot);
  }
  // lanes over j (coalesced weight rows), each warp takes 32 of the chunk's channels
  const int cw0 = warp * 32, cn = min(32, C - c0 - cw0);
  for (int j0 = 0; j0 < m; j0 += 32) {
    const int j = j0 + lane;
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
    if (j < m) {
#pragma unroll 8
      for (int cc = 0; cc < cn; ++cc) {
        const float wv = __ldg(wc + (long long)(c0 + cw0 + cc) * m + j);
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = fmaf(dzs[r * kThreads + cw0 + cc], wv, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < kRows; ++r) atomicAdd(dts + r * m + j, acc[r]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRows * m; i += kThreads) {
    const int r = i / m;
    if (r0 + r < R) atomicAdd(q.dt + ((long long)d * R + r0 + r) * m + (i - r * m), dts[i]);
  }
}

// ---- backward 1b: the cross interaction and the gelu, per row tile: dT -> dHhat, projection / gamma gradients, and
// the BatchNorm-backward partial sums.  grid (nblk).  sm: gate arrays, dT[2][kRows][m], dT0[2][kRows][m]
__global__ void __launch_bounds__(kThreads) ca_gate_bwd_b_kernel(const DmCaGates p, const DmCaGatesGrad q) {
  extern __shared__ __align__(16) float sm[];
  const int m = p.m, R = p.R, r0 = blockIdx.x * kRows, km = kRows * m;
  const GateSmem g = carve_gate(sm, m);
  float* dts = g.pp + 2 * km;
  float* dt0 = dts + 2 * km;
  gate_front(p, g, r0, false);
  const Scalars s = load_scalars(p);
  for (int idx = threadIdx.x; idx < 2 * km; idx += kThreads) {
    const int d = idx / km, rj = idx - d * km, r = rj / m;
    dts[idx] = (r0 + r < R) ? q.dt[((long long)d * R + r0 + r) * m + (rj - r * m)] : 0.0f;
  }
  __syncthreads();
  // cross interaction: T_w = T0_w + sg_w * (T0_h Whw^T + bhw)  =>  dT0_h += sg_w * dT_w Whw, likewise for w
  for (int idx = threadIdx.x; idx < 2 * km; idx += kThreads) {
    const int d = idx / km, rj = idx - d * km, r = rj / m, j = rj - r * m;
    // proj[d] maps T0_d; its output feeds T_{1-d} scaled by sg[1-d]
    const float* wp = d ? p.wp_w2h : p.wp_h2w;
    const float* dto = dts + (1 - d) * km + r * m;
    float acc = 0.0f;
#pragma unroll 8
    for (int k = 0; k < m; ++k) acc = fmaf(dto[k], __ldg(wp + (long long)k * m + j), acc);
    dt0[idx] = dts[idx] + s.sg[1 - d] * acc;
  }
  // the gamma sums (the projection weight / bias gradients come from ca_wgrad_kernel over dT and T0)
  {
    // per-direction totals: threads stride over (r, k), block reduction, one atomic per direction
    __shared__ float red2[2][kThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int d = 0; d < 2; ++d) {
      float part = 0.0f;
      for (int i = threadIdx.x; i < km; i += kThreads) part = fmaf(dts[(1 - d) * km + i], g.pp[d * km + i], part);
      part = dm::warp_sum(part);
      if (lane == 0) red2[d][warp] = part;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      float tot = 0.0f;
      for (int w = 0; w < kThreads / 32; ++w) tot += red2[threadIdx.x][w];
      atomicAdd(q.scal + 2 + (1 - threadIdx.x), tot);     // scal[2] = d sg_h (uses pp[1]), scal[3] = d sg_w (uses pp[0])
    }
  }
  __syncthreads();
  // through the gelu: dHhat, and the BatchNorm-backward partial sums of this block
  for (int idx = threadIdx.x; idx < 2 * km; idx += kThreads) {
    const int d = idx / km, rj = idx - d * km, r = rj / m;
    const float dh = (r0 + r < R) ? dt0[idx] * gelu_grad_exact(g.hh[idx]) : 0.0f;
    dt0[idx] = dh;
    if (r0 + r < R) q.dh[((long long)d * R + r0 + r) * m + (rj - r * m)] = dh;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * m; idx += kThreads) {
    const int d = idx / m, j = idx - d * m;
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (r0 + r < R) {
        const float dh = dt0[d * km + r * m + j];
        const float xh = (p.u[((long long)d * R + r0 + r) * m + j] - g.mean[d * m + j]) * g.rstd[d * m + j];
        s1 += dh; s2 = fmaf(dh, xh, s2);
      }
    float* part = q.part + (((long long)blockIdx.x * 2 + d) * 2) * m;
    part[j] = s1; part[m + j] = s2;
  }
}

// ---- backward 2: BatchNorm backward, then through the first 1x1 convolution.  grid (nblk, ceil(C/256), 2)
__global__ void __launch_bounds__(kThreads) ca_lin1_bwd_kernel(const DmCaGates p, const DmCaGatesGrad q) {
  extern __shared__ __align__(16) float sm[];
  const int d = blockIdx.z, r0 = blockIdx.x * kRows, C = p.C, m = p.m, R = p.R;
  const bool first = blockIdx.y == 0;  // the channel chunk that also owns the per-row-tile side outputs
  float* du = sm;                      // [kRows][m]
  float* s12 = du + kRows * m;         // [2][m]
  __shared__ float scratch[2 * kThreads];
  const float* W1 = d ? p.w1_w : p.w1_h;
  sum_partials(q.part, p.nblk, m, d, scratch, s12, s12 + m);
  if (blockIdx.x == 0 && first)
    for (int j = threadIdx.x; j < m; j += kThreads) {
      (d ? q.g_bn_b_w : q.g_bn_b_h)[j] += s12[j];           // one block per direction owns these
      (d ? q.g_bn_g_w : q.g_bn_g_h)[j] += s12[m + j];
    }
  for (int idx = threadIdx.x; idx < kRows * m; idx += kThreads) {
    const int r = idx / m, j = idx - r * m;
    float v = 0.0f;
    if (r0 + r < R) {
      const float mean = p.stat[d * 2 * m + j], rstd = p.stat[d * 2 * m + m + j];
      const float ga = (d ? p.bn_g_w : p.bn_g_h)[j];
      const float dh = q.dh[((long long)d * R + r0 + r) * m + j];
      const float xh = (p.u[((long long)d * R + r0 + r) * m + j] - mean) * rstd;
      const float inner = p.training ? dh - (s12[j] + xh * s12[m + j]) / (float)R : dh;
      v = ga * rstd * inner;
      if (first) q.du[((long long)d * R + r0 + r) * m + j] = v;        // for the weight-gradient kernel
    }
    du[idx] = v;
  }
  __syncthreads();
  if (first)
    for (int j = threadIdx.x; j < m; j += kThreads) {
      float sb = 0.0f;
#pragma unroll
      for (int r = 0; r < kRows; ++r) sb += du[r * m + j];
      atomicAdd((d ? q.g_b1_w : q.g_b1_h) + j, sb);
    }
  float* dX = d ? q.d_xw : q.d_xh;
  for (int c = blockIdx.y * kThreads + threadIdx.x; c < min(C, (int)(blockIdx.y + 1) * kThreads); c += kThreads) {
    float dx[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) dx[r] = 0.0f;
#pragma unroll 8
    for (int j = 0; j < m; ++j) {
      const float wv = __ldg(W1 + (long long)j * C + c);
#pragma unroll
      for (int r = 0; r < kRows; ++r) dx[r] = fmaf(du[r * m + j], wv, dx[r]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r)
      if (r0 + r < R) dX[(long long)(r0 + r) * C + c] = dx[r];
  }
}

// ---- weight gradients, parallel over outputs (no atomics): out[c*so_c + j*so_j] += sum_r A[r][c] * B[r][j], and
// bias[c] += sum_r A[r][c].  grid (ceil(C/32), ceil(m/8), 2); block = 32 c lanes x 32 row slices.
struct WgArgs {
  const float* A[2]; const float* B[2]; float* out[2]; float* bias[2];
  const float* gate[2];          // optional: the sums are scaled by sigmoid(*gate[d])
  int R, C, m; long long so_c, so_j;
};
__global__ void __launch_bounds__(1024) ca_wgrad_kernel(const WgArgs a) {
  __shared__ float red[32][32][9];
  const int nsl = blockDim.x >> 5;        // row slices: 32 for long reductions, fewer (smaller blocks) for R < 256
  const int d = blockIdx.z, cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, j0 = blockIdx.y * 8;
  const int nj = min(8, a.m - j0);
  const float* A = a.A[d]; const float* B = a.B[d];
  float acc[8], sb = 0.0f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.0f;
  // the accumulate-into-.grad reads go first (one DRAM round trip hidden behind the sums) instead of eight serial
  // load-add-store round trips at the end
  const bool owner = sl == 0 && c < a.C;
  float old[8], oldb = 0.0f;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj)
    old[jj] = (owner && jj < nj) ? __ldcg(a.out[d] + (long long)c * a.so_c + (long long)(j0 + jj) * a.so_j) : 0.0f;
  if (owner && a.bias[d] != nullptr && blockIdx.y == 0) oldb = __ldcg(a.bias[d] + c);
  if (c < a.C)
#pragma unroll 8
    for (int r = sl; r < a.R; r += nsl) {
      const float av = __ldg(A + (long long)r * a.C + c);
      const float* b = B + (long long)r * a.m + j0;
      sb += av;
      if (nj == 8 && (a.m & 3) == 0) {       // 16-byte aligned: j0 is a multiple of 8
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b)), b1 = __ldg(reinterpret_cast<const float4*>(b) + 1);
        acc[0] = fmaf(av, b0.x, acc[0]); acc[1] = fmaf(av, b0.y, acc[1]); acc[2] = fmaf(av, b0.z, acc[2]); acc[3] = fmaf(av, b0.w, acc[3]);
        acc[4] = fmaf(av, b1.x, acc[4]); acc[5] = fmaf(av, b1.y, acc[5]); acc[6] = fmaf(av, b1.z, acc[6]); acc[7] = fmaf(av, b1.w, acc[7]);
      } else {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          if (jj < nj) acc[jj] = fmaf(av, __ldg(b + jj), acc[jj]);
      }
    }
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) red[sl][cl][jj] = acc[jj];
  red[sl][cl][8] = sb;
  __syncthreads();
  const float scale = a.gate[d] != nullptr ? sigmoidf_(a.gate[d][0]) : 1.0f;
  if (owner) {
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      if (jj >= nj) continue;
      float tot = 0.0f;
#pragma unroll
      for (int s2 = 0; s2 < nsl; ++s2) tot += red[s2][cl][jj];
      a.out[d][(long long)c * a.so_c + (long long)(j0 + jj) * a.so_j] = old[jj] + scale * tot;
    }
    if (a.bias[d] != nullptr && blockIdx.y == 0) {
      float tot = 0.0f;
#pragma unroll
      for (int s2 = 0; s2 < nsl; ++s2) tot += red[s2][cl][8];
      a.bias[d][c] = oldb + scale * tot;
    }
  }
}

// ---- backward 3: the four scalar parameters from the accumulated sums (and re-arm the accumulators)
__global__ void ca_scalars_bwd_kernel(const DmCaGates p, const DmCaGatesGrad q) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Scalars s = load_scalars(p);
  const float dka = q.scal[0], dkb = q.scal[1], dsgh = q.scal[2], dsgw = q.scal[3];
  const float inv = 1.0f / (s.S * s.S);
  const float da = (dka * (s.S - s.a) - dkb * s.b) * inv;
  const float db = (dkb * (s.S - s.b) - dka * s.a) * inv;
  q.g_alpha[0] += da * s.a * (1.0f - s.a);
  q.g_beta[0] += db * s.b * (1.0f - s.b);
  q.g_gamma_h[0] += dsgh * s.sg[0] * (1.0f - s.sg[0]);
  q.g_gamma_w[0] += dsgw * s.sg[1] * (1.0f - s.sg[1]);
  q.scal[0] = q.scal[1] = q.scal[2] = q.scal[3] = 0.0f;
}

inline size_t gate_smem_floats(int m, int extra_rows_m) { return (size_t)4 * m + (size_t)(6 + extra_rows_m) * kRows * m; }

int set_smem(const void* fn, size_t bytes, bool& done, size_t& have) {
  if (bytes <= 48 * 1024 || (done && bytes <= have)) return DM_OK;
  if (bytes > 220 * 1024) { dm_set_error("dm_ca_gates: channel count too large for the shared-memory tile"); return DM_ERR_ARG; }
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
  done = true; have = bytes;
  return DM_OK;
}

int check(const DmCaGates* p) {
  if (p->R <= 0 || p->C <= 0 || p->m <= 0 || p->m > kThreads) { dm_set_error("dm_ca_gates: bad sizes"); return DM_ERR_ARG; }
  if (p->nblk != dm::cdiv(p->R, kRows)) { dm_set_error("dm_ca_gates: nblk must be ceil(R / 8)"); return DM_ERR_ARG; }
  return DM_OK;
}

}  // namespace

extern "C" int dm_ca_gates_rows_per_block(void) { return kRows; }

extern "C" int dm_ca_gates_fwd(const DmCaGates* p, void* stream) {
  if (int rc = check(p)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  static bool a1 = false, a2 = false; static size_t h1 = 0, h2 = 0;
  const size_t s1 = (size_t)kRows * p->C * sizeof(float);
  if (int rc = set_smem((const void*)ca_lin1_kernel, s1, a1, h1)) return rc;
  ca_lin1_kernel<<<dim3(p->nblk, 2, dm::cdiv(p->m, kThreads / 32)), kThreads, s1, st>>>(*p);
  DM_CHECK_LAUNCH();
  size_t s2 = gate_smem_floats(p->m, 2) * sizeof(float);
  const size_t slab = (size_t)kThreads * (p->m + 1) * sizeof(float);
  const int use_slab = 0 * (s2 + slab <= 200 * 1024);      // measured slower than the direct reads (the staging loads serialise)
  if (use_slab) s2 += slab;
  if (int rc = set_smem((const void*)ca_gate_fwd_kernel, s2, a2, h2)) return rc;
  ca_gate_fwd_kernel<<<dim3(p->nblk, dm::cdiv(p->C, kThreads)), kThreads, s2, st>>>(*p, use_slab);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

extern "C" int dm_ca_gates_bwd(const DmCaGates* p, const DmCaGatesGrad* q, void* stream) {
  if (int rc = check(p)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  static bool a1 = false, a2 = false; static size_t h1 = 0, h2 = 0;
  const int csplit = dm::cdiv(p->C, kThreads);
  const long long RC = (long long)p->R * p->C, Rm = (long long)p->R * p->m;
  cudaError_t e = cudaMemsetAsync(q->dt, 0, (size_t)(2 * Rm) * sizeof(float), st);
  if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
  const size_t sa = ((size_t)kRows * kThreads + (size_t)kRows * p->m) * sizeof(float);
  if (int rc = set_smem((const void*)ca_gate_bwd_a_kernel, sa, a1, h1)) return rc;
  ca_gate_bwd_a_kernel<<<dim3(p->nblk, csplit, 2), kThreads, sa, st>>>(*p, *q);
  DM_CHECK_LAUNCH();
  const size_t sb = gate_smem_floats(p->m, 4) * sizeof(float);
  if (int rc = set_smem((const void*)ca_gate_bwd_b_kernel, sb, a2, h2)) return rc;
  ca_gate_bwd_b_kernel<<<p->nblk, kThreads, sb, st>>>(*p, *q);
  DM_CHECK_LAUNCH();
  const dim3 wg_grid(dm::cdiv(p->C, 32), dm::cdiv(p->m, 8), 2);
  // row slices per block: about eight rows per thread, so the C = 1536 instance (R = 64 rows, 1152 blocks) runs as one
  // wave of 256-thread blocks instead of four waves of 1024-thread blocks that each add two rows
  int wg_slices = 32;
  while (wg_slices > 1 && wg_slices * 8 > p->R) wg_slices >>= 1;
  const int wg_threads = 32 * wg_slices;
  WgArgs wc = {{q->dz, q->dz + RC}, {p->t, p->t + Rm}, {q->g_wc_h, q->g_wc_w}, {q->g_bc_h, q->g_bc_w}, {nullptr, nullptr},
               p->R, p->C, p->m, (long long)p->m, 1};
  ca_wgrad_kernel<<<wg_grid, wg_threads, 0, st>>>(wc);
  DM_CHECK_LAUNCH();
  // projection gradients: d wp_e[k][j] = sg[1-e] * sum_r dT_{1-e}[r][k] * T0_e[r][j]  (wp_0 = h2w, wp_1 = w2h)
  WgArgs wpj = {{q->dt + Rm, q->dt}, {p->t0, p->t0 + Rm}, {q->g_wp_h2w, q->g_wp_w2h}, {q->g_bp_h2w, q->g_bp_w2h},
                {p->gamma_w, p->gamma_h}, p->R, p->m, p->m, (long long)p->m, 1};
  ca_wgrad_kernel<<<dim3(dm::cdiv(p->m, 32), dm::cdiv(p->m, 8), 2), wg_threads, 0, st>>>(wpj);
  DM_CHECK_LAUNCH();
  const size_t s2 = ((size_t)kRows * p->m + 2 * p->m) * sizeof(float);
  ca_lin1_bwd_kernel<<<dim3(p->nblk, csplit, 2), kThreads, s2, st>>>(*p, *q);
  DM_CHECK_LAUNCH();
  WgArgs w1 = {{p->xh, p->xw}, {q->du, q->du + Rm}, {q->g_w1_h, q->g_w1_w}, {nullptr, nullptr}, {nullptr, nullptr},
               p->R, p->C, p->m, 1, (long long)p->C};
  ca_wgrad_kernel<<<wg_grid, wg_threads, 0, st>>>(w1);
  DM_CHECK_LAUNCH();
  ca_scalars_bwd_kernel<<<1, 32, 0, st>>>(*p, *q);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
