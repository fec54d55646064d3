// DDPM-side kernels (q-sample, weighted loss fwd/bwd, CFG combine + reverse step), weight packing for
// the implicit-GEMM kernels, and the optimizer pass (global grad-norm + clipped AdamW).
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "dm_b200.h"

using dm::bf16;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void dm_set_error(const char* msg) { strncpy(g_err, msg, sizeof(g_err) - 1); g_err[sizeof(g_err) - 1] = 0; }
extern "C" const char* dm_last_error(void) { return g_err; }
extern "C" int dm_version(void) { return 100; }
static long long g_launches = 0;
void dm_count_launch() { ++g_launches; }
extern "C" long long dm_launch_count(void) { return g_launches; }
// Which kernel variant the dispatchers chose: per-name launch counters and the most recent (name, parameter) pair
// (parameter = tile width for the conv kernels, split-K count for the weight-gradient kernels).  Host-side only.
namespace { struct KernelNote { const char* name; long long count; }; KernelNote g_notes[32]; int g_num_notes = 0;
            const char* g_last_name = ""; int g_last_param = 0; }
void dm_note_kernel(const char* name, int param) {
  g_last_name = name; g_last_param = param;
  for (int i = 0; i < g_num_notes; ++i) if (!strcmp(g_notes[i].name, name)) { ++g_notes[i].count; return; }
  if (g_num_notes < 32) { g_notes[g_num_notes].name = name; g_notes[g_num_notes].count = 1; ++g_num_notes; }
}
extern "C" long long dm_kernel_count(const char* name) {
  for (int i = 0; i < g_num_notes; ++i) if (!strcmp(g_notes[i].name, name)) return g_notes[i].count;
  return 0;
}
extern "C" const char* dm_last_kernel(int* param) { if (param) *param = g_last_param; return g_last_name; }

#define ST ((cudaStream_t)stream)

namespace {

inline int grid_for(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  long long cap = (long long)DM_NUM_SMS * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------ packing
struct PackArgs {
  long long tap_off[64];
  int rows, cols, ntaps, c_split, cols_k, tap_major_rows;
  long long s_row, s_col, row_len;
};

// Both kernels run one thread per (row, packed column) and walk the taps, so the accesses to the packed
// matrix (bf16 writes / fp32 reads) are coalesced across the warp and each thread touches ntaps
// consecutive elements of the reference-layout tensor.
__device__ __forceinline__ int pack_col(const PackArgs& A, int col_k) {
  // packed column -> source column, or -1 for a zero pad column
  int col = col_k;
  if (A.c_split > 0) {
    const int s64 = (A.c_split + 63) / 64 * 64;
    if (col_k < s64) { if (col_k >= A.c_split) return -1; }
    else col = col_k - s64 + A.c_split;
  }
  return col < A.cols ? col : -1;
}
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, PackArgs A) {
  const long long total = (long long)A.rows * A.cols_k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col_k = (int)(i % A.cols_k);
    const long long row = i / A.cols_k;
    const int col = pack_col(A, col_k);
    const float* src = w + row * A.s_row + (long long)(col < 0 ? 0 : col) * A.s_col;
    for (int tap = 0; tap < A.ntaps; ++tap) {
      const float v = col < 0 ? 0.0f : __ldg(src + A.tap_off[tap]);
      const long long pk = A.tap_major_rows ? ((long long)tap * A.rows + row) * A.row_len + col_k
                                            : row * A.row_len + (long long)tap * A.cols_k + col_k;
      out[pk] = __float2bfloat16(v);
    }
  }
}
// grad[...] += dwp[...] over the real (row, col, tap) entries; consume != 0 re-zeroes dwp in the same pass.
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(float* __restrict__ dwp, float* __restrict__ grad, PackArgs A,
                                                            int consume) {
  const long long total = (long long)A.rows * A.cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % A.cols);
    const long long row = i / A.cols;
    int col_k = col;
    if (A.c_split > 0 && col >= A.c_split) col_k = col - A.c_split + (A.c_split + 63) / 64 * 64;
    float* dst = grad + row * A.s_row + (long long)col * A.s_col;
    for (int tap = 0; tap < A.ntaps; ++tap) {
      const long long pk = A.tap_major_rows ? ((long long)tap * A.rows + row) * A.row_len + col_k
                                            : row * A.row_len + (long long)tap * A.cols_k + col_k;
      dst[A.tap_off[tap]] += dwp[pk];
      if (consume) dwp[pk] = 0.0f;
    }
  }
}

// ConvTranspose2d(k == s) weights [Cin][Cout][k*k]: the 151 M-parameter up0 layer dominates the generic
// kernels' time (each thread walks 64 taps, lanes 393 KB apart), so its two pack layouts are produced by
// shared-memory tile transposes with contiguous segments on both sides.
//   src S[ci][j], j = co*T + tap  (T = k*k taps, L = Cout*T contiguous)
// K1  forward pack   out[(tap*Cout + co)][ci]      (tap-major rows, K = ci)
__global__ void __launch_bounds__(256) convt_pack_fwd_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cin, int Cout,
                                                              int T, int cin_k) {
  __shared__ float tile[32][33];
  const long long L = (long long)Cout * T;
  const long long j0 = (long long)blockIdx.x * 32;
  const int ci0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r; const long long j = j0 + tx;
    tile[r][tx] = (ci < Cin && j < L) ? __ldg(w + (long long)ci * L + j) : 0.0f;
  }
  __syncthreads();
  for (int c = ty; c < 32; c += 8) {
    const long long j = j0 + c;
    if (j < L && ci0 + tx < cin_k) {
      const int co = (int)(j / T), tap = (int)(j - (long long)co * T);
      out[((long long)tap * Cout + co) * cin_k + ci0 + tx] = __float2bfloat16(ci0 + tx < Cin ? tile[tx][c] : 0.0f);
    }
  }
}
// K2  data-gradient pack   out[ci][tap*Cout + co]    (row_len >= T*Cout)
__global__ void __launch_bounds__(256) convt_pack_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int T,
                                                                long long row_len) {
  extern __shared__ float sm[];                     // [32 co][T + 1]
  const int ci = blockIdx.y, co0 = blockIdx.x * 32;
  const int nco = min(32, Cout - co0);
  const float* src = w + ((long long)ci * Cout + co0) * T;
  for (int i = threadIdx.x; i < nco * T; i += 256) { const int c = i / T, t = i - c * T; sm[c * (T + 1) + t] = __ldg(src + i); }
  __syncthreads();
  bf16* dst = out + (long long)ci * row_len + co0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = warp; t < T; t += 8)
    if (lane < nco) dst[(long long)t * Cout + lane] = __float2bfloat16(sm[lane * (T + 1) + t]);
}
// ------------------------------------------------------------------------------------------ DDPM
__global__ void q_sample_kernel(const float* x, const float* noise, const float* sqrtab, const float* sqrtmab,
                                const long long* ts, bf16* xt, int ldo, int N, int C, int HW) {
  const long long P = (long long)N * HW;
  const int Cv = ldo / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < P * Cv; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % P; const int cv = (int)(i / P);
    const long long n = p / HW, hw = p % HW;
    const long long t = ts[n];
    const float a = sqrtab[t], b = sqrtmab[t];
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      if (c < C) {
        const long long k = (n * C + c) * HW + hw;
        // separate roundings, as the reference's eager mul / mul / add (new_scripy.py:408-411)
        v[j] = __fadd_rn(__fmul_rn(a, x[k]), __fmul_rn(b, noise[k]));
      } else v[j] = 0.f;
    }
    dm::store8(xt + p * ldo + cv * 8, v);
  }
}

struct LossCfg { float hi_t, mid_t, hi_w, mid_w, lo_w, fcw; };

__global__ void __launch_bounds__(256) loss_fwd_kernel(const float* pred, int ldp, const float* noise, const float* mask,
                                                        float* acc, int N, int C, int HW, LossCfg L) {
  __shared__ float sm[32];
  const long long P = (long long)N * HW;
  float s0 = 0.f, s1 = 0.f;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, hw = p % HW;
    float w = 1.f, h = 0.f;
    if (mask) {
      const float m = mask[p];
      w = (m > L.hi_t) ? L.hi_w : ((m > L.mid_t) ? L.mid_w : L.lo_w);     // exact fp32 compares (new_scripy.py:420-424)
      h = (m > L.hi_t) ? 1.f : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      const float pr = pred[p * ldp + c], nz = noise[(n * C + c) * HW + hw];
      const float d = nz - pr;
      s0 += d * d * w;
      s1 += fabsf(pr * h - nz * h);
    }
  }
  s0 = dm::block_sum(s0, sm);
  s1 = dm::block_sum(s1, sm);
  if (threadIdx.x == 0) { atomicAdd(acc, s0); atomicAdd(acc + 1, s1); }
}
__global__ void loss_final_kernel(const float* acc, float* loss, double cnt, float fcw) {
  loss[0] = (float)((double)acc[0] / cnt + (double)fcw * (double)acc[1] / cnt);
}
__global__ void loss_bwd_kernel(const float* pred, int ldp, const float* noise, const float* mask, const float* gout,
                                float* dpred, int lddp, int N, int C, int HW, LossCfg L, float inv_cnt) {
  const long long P = (long long)N * HW;
  const float go = gout[0] * inv_cnt;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long n = p / HW, hw = p % HW;
    float w = 1.f, h = 0.f;
    if (mask) {
      const float m = mask[p];
      w = (m > L.hi_t) ? L.hi_w : ((m > L.mid_t) ? L.mid_w : L.lo_w);
      h = (m > L.hi_t) ? 1.f : 0.f;
    }
    for (int c = 0; c < lddp; ++c) {
      float v = 0.f;
      if (c < C) {
        const float pr = pred[p * ldp + c], nz = noise[(n * C + c) * HW + hw];
        const float e = pr * h - nz * h;
        const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
        v = go * (2.f * w * (pr - nz) + L.fcw * h * sg);
      }
      dpred[p * lddp + c] = v;
    }
  }
}

__global__ void cfg_reverse_kernel(const float* eps, int ldp, const float* x, const float* z, float* x_out, bf16* xt_next,
                                   int ldo, float gw, float a, float b, float s, const float* coef, const float* wvec, int n,
                                   int C, int HW) {
  if (coef != nullptr) { gw = coef[0]; a = coef[1]; b = coef[2]; s = coef[3]; }     // graph-replayable: scalars from device memory
  const long long P = (long long)n * HW;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const long long i = p / HW, hw = p % HW;
    if (wvec != nullptr) gw = wvec[i];               // one guidance scale per trajectory (several scales in one batch)
    for (int c0 = 0; c0 < ldo; c0 += 8) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        if (c < C) {
          const long long k = (i * C + c) * HW + hw;
          const float e1 = eps[p * ldp + c], e2 = eps[(p + P) * ldp + c];
          // eps = (1+w)*eps1 - w*eps2 ; x' = a*(x - eps*b) + s*z, each op rounded separately like the
          // reference's eager kernels (new_scripy.py:470-475)
          const float e = __fsub_rn(__fmul_rn(1.0f + gw, e1), __fmul_rn(gw, e2));
          float r = __fmul_rn(a, __fsub_rn(x[k], __fmul_rn(e, b)));
          if (z) r = __fadd_rn(r, __fmul_rn(s, z[k]));
          x_out[k] = r;
          v[j] = r;
        } else v[j] = 0.f;
      }
      if (xt_next) {
        dm::store8(xt_next + p * ldo + c0, v);
        dm::store8(xt_next + (p + P) * ldo + c0, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ optimizer
// Deterministic: per-block partial sums, combined in a fixed order (double precision) by the last block to arrive.  The
// clip coefficient derived from this norm scales the whole update, so under data parallelism it must come out BIT-IDENTICAL
// on every rank from bit-identical all-reduced gradients (an atomicAdd per block did not: the ranks' parameters drifted
// apart in the last bits step by step).  out[0] is overwritten (no zero-fill needed); ws = [kSumsqSlots partials][counter].
constexpr int kSumsqSlots = DM_NUM_SMS * 16;          // = the grid cap of grid_for()
__global__ void __launch_bounds__(256) sumsq_kernel(const float* g, long long n, float* out, float* ws) {
  __shared__ float sm[32];
  __shared__ bool last;
  float s = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
  s = dm::block_sum(s, sm);
  unsigned* counter = reinterpret_cast<unsigned*>(ws + kSumsqSlots);          // fixed slot: grids of different sizes share ws
  if (threadIdx.x == 0) {
    ws[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  __shared__ double sd[256];
  double a = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += (double)__ldcg(ws + i);
  sd[threadIdx.x] = a;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if ((int)threadIdx.x < k) sd[threadIdx.x] += sd[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = (float)sd[0]; *counter = 0u; }
}

// torch.optim.AdamW semantics (decoupled decay, bias-corrected), gradient pre-scaled by the
// clip_grad_norm_ coefficient min(1, max_norm / (norm + 1e-6)) (new_scripy.py:798-801).
__global__ void __launch_bounds__(256) adamw_kernel(float* p, const float* g, float* m, float* v, long long n, float lr,
                                                     float b1, float b2, float eps, float wd, float bc1, float bc2,
                                                     const float* gnorm_sq, float max_norm) {
  float clip = 1.f;
  if (gnorm_sq != nullptr && max_norm > 0.f) {
    const float c = max_norm / (sqrtf(gnorm_sq[0]) + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const float step = lr / bc1, rs2 = 1.0f / sqrtf(bc2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    pi -= step * mi / (sqrtf(vi) * rs2 + eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// Same update on float4 vectors, additionally writing a bf16 copy of the new parameters: for convolution weights
// stored in the GEMM's own [Cout][tap][Cin] order that copy IS the forward weight pack, so no pack kernel runs.
__global__ void __launch_bounds__(256) adamw_bf16_kernel(float4* p, const float4* g, float4* m, float4* v, uint2* pb,
                                                          long long n4, float lr, float b1, float b2, float eps, float wd,
                                                          float bc1, float bc2, const float* gnorm_sq, float max_norm) {
  float clip = 1.f;
  if (gnorm_sq != nullptr && max_norm > 0.f) {
    const float c = max_norm / (sqrtf(gnorm_sq[0]) + 1e-6f);
    clip = c < 1.f ? c : 1.f;
  }
  const float step = lr / bc1, rs2 = 1.0f / sqrtf(bc2), decay = 1.f - lr * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = g[i]; float4 p4 = p[i], m4 = m[i], v4 = v[i];
    float* pp = reinterpret_cast<float*>(&p4); float* mm = reinterpret_cast<float*>(&m4); float* vv = reinterpret_cast<float*>(&v4);
    const float* gg = reinterpret_cast<const float*>(&g4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = gg[j] * clip;
      float pi = pp[j] * decay;
      const float mi = mm[j] + (1.f - b1) * (gi - mm[j]);
      const float vi = b2 * vv[j] + (1.f - b2) * gi * gi;
      pi -= step * mi / (sqrtf(vi) * rs2 + eps);
      pp[j] = pi; mm[j] = mi; vv[j] = vi;
    }
    p[i] = p4; m[i] = m4; v[i] = v4;
    pb[i] = make_uint2(dm::pack2(p4.x, p4.y), dm::pack2(p4.z, p4.w));
  }
}

// Data-gradient packs from the bf16 forward pack of a GEMM-native weight: per filter tap a [Cout x Cin] ->
// [Cin x Cout] transpose through a 64x64 shared-memory tile (128-byte contiguous segments on both sides), the tap
// landing at dst_off[tap] (180-degree rotation for stride 1, the four parity phases for 4x4/s2).
//   src[co][tap][ci] (pitch ntaps*cin)   ->   dst[ci * dst_pitch + dst_off[tap] + co], co < ck (zero beyond Cout)
struct TransArgs { long long dst_off[16]; };
__global__ void __launch_bounds__(256) pack_transpose_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, int cout,
                                                              int cin, int ntaps, long long dst_pitch, TransArgs A) {
  // cin is a multiple of 64 (GEMM-native weights) and dst rows are ceil64(cout) wide: 4-byte accesses throughout
  __shared__ __align__(4) bf16 tile[64][66];
  const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64, tap = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 64; r += 8) {
    const int co = co0 + r;
    uint32_t v = 0u;
    if (co < cout) v = *reinterpret_cast<const uint32_t*>(src + ((long long)co * ntaps + tap) * cin + ci0 + 2 * tx);
    *reinterpret_cast<uint32_t*>(&tile[r][2 * tx]) = v;
  }
  __syncthreads();
  const unsigned short* t16 = reinterpret_cast<const unsigned short*>(&tile[0][0]);
  for (int r = ty; r < 64; r += 8) {
    const uint32_t lo = t16[(2 * tx) * 66 + r], hi = t16[(2 * tx + 1) * 66 + r];
    *reinterpret_cast<uint32_t*>(dst + (long long)(ci0 + r) * dst_pitch + A.dst_off[tap] + co0 + 2 * tx) = lo | (hi << 16);
  }
}

// Input pipeline tail on the device (CrackDataset.__getitem__ + the torchvision transforms, new_scripy.py:516-551,683-688)
// for a cached uint8 batch [B][H][W][3] already decoded and resized: RandomHorizontalFlip (decision per sample from
// the host RNG), ToTensor (u8 / 255), Normalize ((v - mean) / std, separately rounded like the eager ops) -> x fp32
// NCHW, and the attention mask: `low` everywhere, `mid` in the lower half, `high` inside the scaled bounding box
// [ymin, ymax) x [xmin, xmax) (the reference does NOT flip the mask with the image).
__global__ void __launch_bounds__(256) prep_batch_kernel(const unsigned char* __restrict__ img, const int* __restrict__ flip,
                                                          const int* __restrict__ box, float* __restrict__ x,
                                                          float* __restrict__ mask, int B, int H, int W, float mean, float stdv,
                                                          float low, float mid, float high) {
  const long long total = (long long)B * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W), h = (int)((i / W) % H), b = (int)(i / ((long long)W * H));
    const int ws = flip[b] ? W - 1 - w : w;
    const unsigned char* px = img + (((long long)b * H + h) * W + ws) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      x[(((long long)b * 3 + c) * H + h) * W + w] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[c], 255.0f), mean), stdv);
    const int* bx = box + 4 * b;                       // xmin, ymin, xmax, ymax (scaled, clamped)
    float m = h >= H / 2 ? mid : low;
    if (h >= bx[1] && h < bx[3] && w >= bx[0] && w < bx[2]) m = high;
    mask[i] = m;
  }
}

// Per-image-pair quality metrics of ImageMetrics.calc_ssim / calc_psnr (new_scripy.py:1189-1251): global-statistics SSIM
// and PSNR on images mapped to [0,1] ("(x+1)/2 if x.min() < 0", decided per image).  One block per pair: pass 1 the two
// minima, pass 2 six sums in double precision.  out[n] = (ssim, psnr); psnr = +inf for identical images.
__global__ void __launch_bounds__(1024) image_metrics_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              float* __restrict__ out, long long elems) {
  __shared__ float smin[2][32];
  __shared__ double ssum[6][32];
  const float* pa = a + (long long)blockIdx.x * elems;
  const float* pb = b + (long long)blockIdx.x * elems;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ma = 3.4e38f, mb = 3.4e38f;
  for (long long i = threadIdx.x; i < elems; i += blockDim.x) { ma = fminf(ma, pa[i]); mb = fminf(mb, pb[i]); }
  for (int o = 16; o > 0; o >>= 1) { ma = fminf(ma, __shfl_xor_sync(0xffffffffu, ma, o)); mb = fminf(mb, __shfl_xor_sync(0xffffffffu, mb, o)); }
  if (lane == 0) { smin[0][warp] = ma; smin[1][warp] = mb; }
  __syncthreads();
  ma = smin[0][0]; mb = smin[1][0];
  for (int w = 1; w < 32; ++w) { ma = fminf(ma, smin[0][w]); mb = fminf(mb, smin[1][w]); }
  const bool ca = ma < 0.f, cb = mb < 0.f;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (long long i = threadIdx.x; i < elems; i += blockDim.x) {
    float x = pa[i], y = pb[i];
    if (ca) x = (x + 1.f) / 2.f;
    if (cb) y = (y + 1.f) / 2.f;
    const double dx = x, dy = y, df = (double)(x - y);
    s[0] += dx; s[1] += dy; s[2] += dx * dx; s[3] += dy * dy; s[4] += dx * dy; s[5] += df * df;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = s[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) ssum[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[6];
    for (int k = 0; k < 6; ++k) { t[k] = 0; for (int w = 0; w < 32; ++w) t[k] += ssum[k][w]; }
    const double n = (double)elems, mu1 = t[0] / n, mu2 = t[1] / n;
    const double v1 = t[2] / n - mu1 * mu1, v2 = t[3] / n - mu2 * mu2, c12 = t[4] / n - mu1 * mu2, mse = t[5] / n;
    const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
    out[2 * blockIdx.x] = (float)(((2 * mu1 * mu2 + C1) * (2 * c12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (v1 + v2 + C2)));
    out[2 * blockIdx.x + 1] = mse == 0.0 ? __int_as_float(0x7f800000) : (float)(20.0 * log10(1.0 / sqrt(mse)));
  }
}

}  // namespace

static void fill_pack(PackArgs& A, int rows, int cols, int ntaps, const long long* tap_off, long long s_row, long long s_col,
                      int c_split, int cols_k, long long row_len, int tap_major_rows) {
  memset(&A, 0, sizeof A);
  A.rows = rows; A.cols = cols; A.ntaps = ntaps; A.s_row = s_row; A.s_col = s_col; A.c_split = c_split;
  A.cols_k = cols_k; A.row_len = row_len; A.tap_major_rows = tap_major_rows;
  for (int i = 0; i < ntaps && i < 64; ++i) A.tap_off[i] = tap_off[i];
}

extern "C" int dm_pack_weight(const float* w, void* out, int rows, int cols, int ntaps, const long long* tap_off_host,
                              long long s_row, long long s_col, int c_split, int cols_k, long long row_len,
                              int tap_major_rows, void* stream) {
  if (ntaps < 1 || ntaps > 64) { dm_set_error("dm_pack_weight: 1..64 taps"); return DM_ERR_ARG; }
  PackArgs A; fill_pack(A, rows, cols, ntaps, tap_off_host, s_row, s_col, c_split, cols_k, row_len, tap_major_rows);
  bool ident = true;
  for (int i = 0; i < ntaps; ++i) ident = ident && tap_off_host[i] == i;
  // ConvTranspose2d weight [Cin][Cout][T]: tiled transposes (see convt_pack_*_kernel)
  if (ident && c_split == 0 && tap_major_rows && s_row == ntaps && s_col == (long long)rows * ntaps && row_len == cols_k &&
      (long long)rows * cols >= 4096) {
    dim3 grid(dm::cdiv((long long)rows * ntaps, 32), dm::cdiv(cols_k, 32));
    convt_pack_fwd_kernel<<<grid, 256, 0, ST>>>(w, (bf16*)out, cols, rows, ntaps, cols_k);
    DM_CHECK_LAUNCH();
    return DM_OK;
  }
  if (ident && c_split == 0 && !tap_major_rows && s_col == ntaps && s_row == (long long)cols * ntaps && cols_k == cols &&
      ntaps > 16 && (long long)rows * cols >= 4096) {
    if (row_len > (long long)ntaps * cols_k) {
      cudaError_t e = cudaMemsetAsync(out, 0, (size_t)rows * row_len * 2, ST);
      if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
    }
    dim3 grid(dm::cdiv(cols, 32), rows);
    convt_pack_dgrad_kernel<<<grid, 256, (size_t)32 * (ntaps + 1) * sizeof(float), ST>>>(w, (bf16*)out, cols, ntaps, row_len);
    DM_CHECK_LAUNCH();
    return DM_OK;
  }
  if (!tap_major_rows && row_len > (long long)ntaps * cols_k) {      // zero the row tails the kernel never writes
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)rows * row_len * 2, ST);
    if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
  }
  pack_weight_kernel<<<grid_for((long long)rows * cols_k), 256, 0, ST>>>(w, (bf16*)out, A);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_unpack_wgrad(float* dwp, float* grad, int rows, int cols, int ntaps, const long long* tap_off_host,
                               long long s_row, long long s_col, int c_split, int cols_k, long long row_len,
                               int tap_major_rows, int consume, void* stream) {
  if (ntaps < 1 || ntaps > 64) { dm_set_error("dm_unpack_wgrad: 1..64 taps"); return DM_ERR_ARG; }
  PackArgs A; fill_pack(A, rows, cols, ntaps, tap_off_host, s_row, s_col, c_split, cols_k, row_len, tap_major_rows);
  unpack_wgrad_kernel<<<grid_for((long long)rows * cols), 256, 0, ST>>>(dwp, grad, A, consume);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

extern "C" int dm_q_sample(const float* x, const float* noise, const float* sqrtab, const float* sqrtmab, const long long* ts,
                           void* xt, int ldo, int N, int C, int H, int W, void* stream) {
  if (ldo & 7) { dm_set_error("dm_q_sample: pitch must be a multiple of 8"); return DM_ERR_ARG; }
  q_sample_kernel<<<grid_for((long long)N * H * W * (ldo / 8)), 256, 0, ST>>>(x, noise, sqrtab, sqrtmab, ts, (bf16*)xt, ldo, N, C, H * W);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_ddpm_loss_fwd(const float* pred, int ldp, const float* noise, const float* mask, float* loss, float* scratch,
                                int N, int C, int H, int W, float hi_t, float mid_t, float hi_w, float mid_w, float lo_w,
                                float fcw, void* stream) {
  cudaError_t e = cudaMemsetAsync(scratch, 0, 2 * sizeof(float), ST);
  if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
  LossCfg L{hi_t, mid_t, hi_w, mid_w, lo_w, mask ? fcw : 0.f};
  loss_fwd_kernel<<<grid_for((long long)N * H * W), 256, 0, ST>>>(pred, ldp, noise, mask, scratch, N, C, H * W, L);
  DM_CHECK_LAUNCH();
  loss_final_kernel<<<1, 1, 0, ST>>>(scratch, loss, (double)N * C * H * W, L.fcw);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_ddpm_loss_bwd(const float* pred, int ldp, const float* noise, const float* mask, const float* gout,
                                void* dpred, int lddp, int N, int C, int H, int W, float hi_t, float mid_t, float hi_w,
                                float mid_w, float lo_w, float fcw, void* stream) {
  LossCfg L{hi_t, mid_t, hi_w, mid_w, lo_w, mask ? fcw : 0.f};
  loss_bwd_kernel<<<grid_for((long long)N * H * W), 256, 0, ST>>>(pred, ldp, noise, mask, gout, (float*)dpred, lddp, N, C, H * W, L,
                                                                 (float)(1.0 / ((double)N * C * H * W)));
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_cfg_reverse_step(const float* eps, int ldp, const float* x, const float* z, float* x_out, void* xt_next,
                                   int ldo, float guide_w, float oneover_sqrta, float mab_over_sqrtmab, float sqrt_beta,
                                   int n, int C, int H, int W, void* stream) {
  if (ldo & 7) { dm_set_error("dm_cfg_reverse_step: pitch must be a multiple of 8"); return DM_ERR_ARG; }
  cfg_reverse_kernel<<<grid_for((long long)n * H * W), 256, 0, ST>>>(eps, ldp, x, z, x_out, (bf16*)xt_next, ldo, guide_w,
                                                                    oneover_sqrta, mab_over_sqrtmab, sqrt_beta, nullptr, nullptr, n, C, H * W);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_cfg_reverse_step_dev(const float* eps, int ldp, const float* x, const float* z, float* x_out, void* xt_next,
                                       int ldo, const float* coef4, int n, int C, int H, int W, void* stream) {
  if (ldo & 7) { dm_set_error("dm_cfg_reverse_step_dev: pitch must be a multiple of 8"); return DM_ERR_ARG; }
  if (coef4 == nullptr) { dm_set_error("dm_cfg_reverse_step_dev: coefficient buffer required"); return DM_ERR_ARG; }
  cfg_reverse_kernel<<<grid_for((long long)n * H * W), 256, 0, ST>>>(eps, ldp, x, z, x_out, (bf16*)xt_next, ldo, 0.f, 0.f, 0.f, 0.f,
                                                                    coef4, nullptr, n, C, H * W);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
/* same with one guidance scale per trajectory: w[n] replaces coef4[0] (several --guide_scales as ONE trajectory batch) */
extern "C" int dm_cfg_reverse_step_w(const float* eps, int ldp, const float* x, const float* z, float* x_out, void* xt_next,
                                     int ldo, const float* coef4, const float* w, int n, int C, int H, int W, void* stream) {
  if (ldo & 7) { dm_set_error("dm_cfg_reverse_step_w: pitch must be a multiple of 8"); return DM_ERR_ARG; }
  if (coef4 == nullptr || w == nullptr) { dm_set_error("dm_cfg_reverse_step_w: coefficient and scale buffers required"); return DM_ERR_ARG; }
  cfg_reverse_kernel<<<grid_for((long long)n * H * W), 256, 0, ST>>>(eps, ldp, x, z, x_out, (bf16*)xt_next, ldo, 0.f, 0.f, 0.f, 0.f,
                                                                    coef4, w, n, C, H * W);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

extern "C" int dm_sumsq(const float* g, long long n, float* out, void* stream) {
  static float* ws = nullptr;          // partial sums + arrival counter (one device per process, like the reduction workspace)
  const int grid = grid_for(n / 4 + 1);
  if (ws == nullptr) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ST, &cap);
    if (cap != cudaStreamCaptureStatusNone) { dm_set_error("dm_sumsq: first call inside a stream capture (scratch not allocated yet)"); return DM_ERR_ARG; }
    const size_t bytes = ((size_t)kSumsqSlots + 1) * sizeof(float);
    if (cudaMalloc(&ws, bytes) != cudaSuccess || cudaMemset(ws, 0, bytes) != cudaSuccess) { dm_set_error("dm_sumsq: scratch allocation failed"); return DM_ERR_CUDA; }
  }
  sumsq_kernel<<<grid, 256, 0, ST>>>(g, n, out, ws);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
extern "C" int dm_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                        float eps, float wd, float bc1, float bc2, const float* gnorm_sq, float max_norm, void* stream) {
  adamw_kernel<<<grid_for(n), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2, gnorm_sq, max_norm);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

/* dm_adamw plus a bf16 shadow copy of the updated parameters (n a multiple of 4, all pointers 16-byte aligned) */
extern "C" int dm_adamw_bf16(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                             float beta2, float eps, float wd, float bc1, float bc2, const float* gnorm_sq, float max_norm,
                             void* stream) {
  if (n & 3) { dm_set_error("dm_adamw_bf16: element count must be a multiple of 4"); return DM_ERR_ARG; }
  adamw_bf16_kernel<<<grid_for(n / 4), 256, 0, ST>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)p_bf16, n / 4,
                                                     lr, beta1, beta2, eps, wd, bc1, bc2, gnorm_sq, max_norm);
  DM_CHECK_LAUNCH();
  return DM_OK;
}


/* bf16 [Cout][ntaps][Cin] -> dst[ci*dst_pitch + dst_off[tap] + co] for co < ck = ceil64(Cout) (zeros beyond Cout) */
extern "C" int dm_pack_transpose(const void* src, void* dst, int cout, int cin, int ntaps, const long long* dst_off_host,
                                 long long dst_pitch, void* stream) {
  if (ntaps < 1 || ntaps > 16) { dm_set_error("dm_pack_transpose: 1..16 taps"); return DM_ERR_ARG; }
  TransArgs A;
  for (int i = 0; i < 16; ++i) A.dst_off[i] = i < ntaps ? dst_off_host[i] : 0;
  if (cin % 64) { dm_set_error("dm_pack_transpose: Cin must be a multiple of 64"); return DM_ERR_ARG; }
  dim3 grid(cin / 64, dm::cdiv(cout, 64), ntaps);
  pack_transpose_kernel<<<grid, 256, 0, ST>>>((const bf16*)src, (bf16*)dst, cout, cin, ntaps, dst_pitch, A);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

/* cached uint8 batch -> normalised fp32 NCHW images + attention masks (new_scripy.py:516-551,683-688) */
extern "C" int dm_prep_batch(const void* img_u8, const int* flip, const int* box, float* x, float* mask, int B, int H, int W,
                             float mean, float stdv, float low, float mid, float high, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return DM_OK;
  prep_batch_kernel<<<grid_for((long long)B * H * W), 256, 0, ST>>>((const unsigned char*)img_u8, flip, box, x, mask, B, H, W,
                                                                    mean, stdv, low, mid, high);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

/* SSIM (global statistics) and PSNR of N image pairs, ImageMetrics.calc_ssim / calc_psnr (new_scripy.py:1189-1251) */
extern "C" int dm_image_metrics(const float* a, const float* b, float* out, int N, long long elems, void* stream) {
  if (N <= 0 || elems <= 0) return DM_OK;
  image_metrics_kernel<<<N, 1024, 0, ST>>>(a, b, out, elems);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
