// Skinny GEMM out[m][n] = sum_k A[m][k] * W[n][k] for M <= 16 rows against a large K-major weight matrix: the data
// gradient of up0 = ConvTranspose2d(8F, 8F, 8, 8) on the 2x2 bottleneck (new_scripy.py:297-301), where M = batch * 2 * 2 = 16
// pixels, N = 1536 input channels and K = 64 taps * 1536 output channels = 98 304.  The work is one pass over the 302 MB
// bf16 weight pack (HBM bound, 4.8 GFLOP); as an implicit-GEMM conv it has 1 x 9 output tiles, i.e. 9 busy SMs.  Here the
// K dimension is split over the grid instead: a block owns 64 weight rows x a 2048-wide K slice, stages the 16 x 2048 slice
// of A in shared memory once, and each warp streams 8 weight rows straight from global memory into mma.sync.m16n8k16
// B fragments (M = 16 is exactly one fragment; the tensor pipe is idle 7/8 of the time and irrelevant, the weight stream is
// the bound).  The contraction index is permuted inside every 32-wide block so that a lane's B fragments for two
// consecutive MMAs are ONE contiguous 16-byte load, and its A fragments two conflict-free 16-byte shared loads.
// Partial sums per K slice go to a workspace; a second kernel folds them in a fixed order (deterministic) and rounds to bf16.
#include "common.cuh"
#include "dm_b200.h"

namespace {

constexpr int kKc = 2048;             // K slice per block
constexpr int kRowsPerBlock = 64;     // weight rows per block (8 warps x 8)
constexpr int kPitch = kKc * 2 + 64;  // bytes per staged A row: +64 B skews rows by 4 bank groups -> conflict-free LDS.128

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256) skinny_gemm_kernel(const dm::bf16* __restrict__ A, long long lda,
                                                          const dm::bf16* __restrict__ W, long long ldw,
                                                          float* __restrict__ part, int M, int N, int K) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int k0 = blockIdx.y * kKc, kc = min(kKc, K - k0);
  // stage A[0..16)[k0 .. k0+kc) (rows >= M are zero)
  for (int v = threadIdx.x; v < 16 * (kKc / 8); v += blockDim.x) {
    const int r = v / (kKc / 8), c = v % (kKc / 8);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r < M && c * 8 < kc) val = dm::ldg16(A + (long long)r * lda + k0 + c * 8);
    *reinterpret_cast<uint4*>(smem + (size_t)r * kPitch + c * 16) = val;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int n0 = blockIdx.x * kRowsPerBlock + warp * 8;
  if (n0 >= N) return;
  const int nrow = min(n0 + g, N - 1);                       // clamp: rows past N are computed but never stored
  const dm::bf16* wrow = W + (long long)nrow * ldw + k0 + q * 8;
  const unsigned char* a_lo = smem + (size_t)g * kPitch + q * 16;
  const unsigned char* a_hi = a_lo + 8 * (size_t)kPitch;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int nblk = kc / 32;
  int b = 0;
  for (; b + 8 <= nblk; b += 8) {
    uint4 w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = dm::ldg16(wrow + (b + u) * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint4 x = *reinterpret_cast<const uint4*>(a_lo + (b + u) * 64);
      const uint4 y = *reinterpret_cast<const uint4*>(a_hi + (b + u) * 64);
      mma16816(acc, x.x, y.x, x.y, y.y, w[u].x, w[u].y);
      mma16816(acc, x.z, y.z, x.w, y.w, w[u].z, w[u].w);
    }
  }
  for (; b < nblk; ++b) {
    const uint4 w = dm::ldg16(wrow + b * 32);
    const uint4 x = *reinterpret_cast<const uint4*>(a_lo + b * 64);
    const uint4 y = *reinterpret_cast<const uint4*>(a_hi + b * 64);
    mma16816(acc, x.x, y.x, x.y, y.y, w.x, w.y);
    mma16816(acc, x.z, y.z, x.w, y.w, w.z, w.w);
  }
  // accumulator layout: rows g / g+8 of A, columns n0 + 2q, n0 + 2q + 1
  float* p = part + ((long long)blockIdx.y * 16) * N;
  const int n = n0 + 2 * q;
  if (n < N) { p[(long long)g * N + n] = acc[0]; p[(long long)(g + 8) * N + n] = acc[2]; }
  if (n + 1 < N) { p[(long long)g * N + n + 1] = acc[1]; p[(long long)(g + 8) * N + n + 1] = acc[3]; }
}

__global__ void __launch_bounds__(256) skinny_fold_kernel(const float* __restrict__ part, int slices, dm::bf16* __restrict__ out,
                                                          int ldo, int M, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float v = 0.f;
  for (int s0 = 0; s0 < slices; s0 += 8) {          // eight partial rows requested at a time, folded in slice order
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = s0 + k < slices ? __ldg(part + ((long long)(s0 + k) * 16 + m) * N + n) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += t[k];
  }
  out[(long long)m * ldo + n] = __float2bfloat16(v);
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" long long dm_skinny_gemm_scratch(int N, int K) { return (long long)((K + kKc - 1) / kKc) * 16 * N; }

extern "C" int dm_skinny_gemm(const void* A, long long lda, const void* W, long long ldw, void* out, int ldo, float* scratch,
                              int M, int N, int K, void* stream) {
  if (M < 1 || M > 16) { dm_set_error("dm_skinny_gemm: M must be 1..16"); return DM_ERR_ARG; }
  if (K % 32 || (lda & 7) || (ldw & 7) || K <= 0 || N <= 0) {
    dm_set_error("dm_skinny_gemm: K must be a positive multiple of 32 and the row pitches multiples of 8");
    return DM_ERR_ARG;
  }
  static bool attr = false;
  const size_t smem = 16 * (size_t)kPitch;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
    attr = true;
  }
  const int slices = (K + kKc - 1) / kKc;
  dim3 grid((N + kRowsPerBlock - 1) / kRowsPerBlock, slices);
  dm_note_kernel("skinny_gemm", (int)grid.x);
  skinny_gemm_kernel<<<grid, 256, smem, ST>>>((const dm::bf16*)A, lda, (const dm::bf16*)W, ldw, scratch, M, N, K);
  DM_CHECK_LAUNCH();
  skinny_fold_kernel<<<(M * N + 255) / 256, 256, 0, ST>>>(scratch, slices, (dm::bf16*)out, ldo, M, N);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
