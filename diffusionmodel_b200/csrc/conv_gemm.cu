// tcgen05 / TMEM / TMA implicit-GEMM kernels for the U-Net convolutions (sm_100a only).
//
//   kernel A  conv_gemm_kernel   D[pixels, Cout] = im2col(X)[pixels, taps*Cin] * Wp[Cout, taps*Cin]^T
//             forward conv (3x3 s1, 1x1, 4x4 s2 through four parity views), data-gradient (same
//             kernel with the flipped/transposed weight pack), conv-transpose k==s (scatter epilogue).
//             M tile = 128 output pixels laid out as a (bn x bh x bw) patch, so one 4-D TMA box per
//             filter tap (shifted by the tap offset, out-of-bounds zero-filled = the conv padding)
//             lands in shared memory already in the K-major SWIZZLE_128B layout tcgen05.mma wants.
//   kernel B  wgrad_gemm_kernel  dWp[Cout, tap, Cin] += dY[pixels, Cout]^T * im2col(X)[pixels, tap*Cin]
//             both operands MN-major (channels contiguous), K = pixels, split-K with fp32 red.add.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / BN partial statistics -> bf16 -> global).
// Pipelines: smem ring full/empty mbarriers (TMA <-> MMA), double-buffered TMEM accumulator
// full/empty mbarriers (MMA <-> epilogue), static persistent tile schedule (grid <= #SMs).
//
// Replaces (reference, all through cuDNN): nn.Conv2d call sites new_scripy.py:166,169,184,188,217,222,
// 225,229,243,311,314 and nn.ConvTranspose2d new_scripy.py:298 / MNIST_script.py:88,141.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "dm_b200.h"

namespace {

using dm::bf16;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;        // bf16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kAStage = kBlockM * kBlockK * 2;       // 16 KB
constexpr int kThreads = 192;        // weight-gradient kernels: TMA warp, MMA warp, 4 epilogue warps
constexpr int kConvThreads = 320;    // conv kernel: TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kSmemBudget = 227 * 1024;
constexpr int kAuxBytes = 512;                       // barriers + TMEM slot (the BN-statistics slab follows, sized per launch)
constexpr int kMaxStatBytes = 56 * 1024;             // fused statistics: 1536 output channels + one tile width of padding
constexpr int kMaxStages = 8;
// halo kernel: an 8 x 16 pixel tile's 3x3 neighbourhood = 10 x 18 pixels x 64 channels, loaded ONCE per
// 64-channel chunk; the nine filter taps are nine shifted views of it (descriptor start + (dy*10+dx) rows,
// 8-row groups 10 rows = 1280 B apart).
constexpr int kHaloW = 10, kHaloH = 18;
constexpr int kHaloBytes = kHaloW * kHaloH * 128;             // 23040
constexpr int kHaloStage = (kHaloBytes + 1023) / 1024 * 1024; // 23552: keeps every patch 1024-aligned
constexpr int kHaloAStages = 2;

struct TapInfo { int8_t dy, dx, map, pad_; };

struct ConvParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  TapInfo taps[16];
  int num_taps, chunks0, chunks1, dual;
  int log_bw, log_bh, log_bn;
  int tiles_w, tiles_h, tiles_b, m_tiles, n_tiles, block_n;
  int N, H, W;
  int stages, b_stage_bytes;
  int stat_c;                   // channels covered by the statistics slab (n_tiles * block_n), 0 = no statistics
  int stat_rows;                // rows of the caller's statistics buffer (>= grid; the surplus is zero-filled)
  void* out;
  int out_f32;
  long long sN, sH, sW;
  int Cout, ldc_pad;
  int convt_k;
  long long sKy, sKx;
  const float* bias;
  const float* scale;           // optional per-channel multiplier applied before the bias (folded eval-mode BatchNorm)
  int act;                      // activation fused after scale/bias: 0 none, 1 GELU, 2 ReLU
  float* stats;
  int stats_ld;
  unsigned idesc;
  unsigned long long desc_hi;   // upper 32 bits of the smem descriptors (SBO / version / layout)
  // halo kernel (3x3, stride 1): ring of input patches + ring of weight tiles
  int a_stages, base_off_mode;
  CUtensorMap tmB2;             // CTA-pair kernel: half-height weight box (block_n/2 rows)
  int pair_tiles;               // CTA-pair kernels: number of (two M tiles) x (N tile) work items
  int pair_m;                   // CTA-pair kernels: M-tile pairs per phase = ceil(m_tiles / 2)
  // generic kernels: `phases` > 1 runs that many independent convs of the same geometry in ONE launch (the four output
  // parities of the stride-2 data gradient): phase f uses taps[f * num_taps ..], weight rows f * phase_rows .. and writes
  // at out + (f >> 1) * out_ph + (f & 1) * out_pw
  int phases, phase_rows;
  long long out_ph, out_pw;
  int ablate;                   // dev: bit0 no MMA issue, bit1 no TMA loads, bit2 no epilogue work (timing decomposition)
  unsigned long long desc_hi_halo;
};

struct WgradParams {
  CUtensorMap tmX[4];
  CUtensorMap tmDY;
  TapInfo taps[16];
  int num_taps, chunks0, chunks1, dual;
  int log_bw, log_bh, log_bn;       // 64-pixel patch
  int tiles_w, tiles_h, tiles_b, patches;
  int co_tiles, ci_tiles, block_n, splits, patches_per_split;
  int Cout, cin_k;                  // cin_k = padded K columns per tap in dWp
  int stages, b_stage_bytes;
  float* dwp;
  unsigned idesc;
  unsigned long long desc_hi_a, desc_hi_b;
  unsigned lbo_a, lbo_b;
};

// kernel C (wgrad2): the transposed product.  M side = im2col(X): four [64 px][64 ch] boxes, each its own
// (filter tap, 64-channel chunk) -> two M=128 accumulators per CTA that share every dY (N side) tile, so
// a K-step moves (4 + nb) x 8 KB for 2 x 128 x N x 64 MACs.  Layers whose Cout is not a multiple of 128
// (192, 48, 3) waste no MMA rows this way.
struct Wgrad2Params {
  CUtensorMap tmX[4];
  CUtensorMap tmDY;
  TapInfo taps[16];
  int num_taps, chunks0, chunks1, dual;
  int log_bw, log_bh, log_bn;       // 64-pixel patch
  int tiles_w, tiles_h, tiles_b, patches;
  int m_tiles, n_tiles, block_n, nb, splits, patches_per_split, total_boxes;
  int Cout, cin_k, C0, C1;
  int stages;
  float* dwp;
  unsigned idesc;
  unsigned long long desc_hi;
  unsigned lbo;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
  }
  printf("dm_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's shared memory, the transaction bytes are
// signalled on the LEADER CTA's mbarrier (a shared::cluster address).
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {     // arrives on the barrier at this offset in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_alloc2(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// One lane of a fully converged warp.  The producer / MMA warps run their loops warp-uniformly and only the
// issue statements sit under elect_one(): inside an `if (lane == 0)` region the compiler must wrap every
// uniform-datapath instruction (UTCHMMA, UTMALDG, UTCBAR) in an ELECT / BRA.U.ANY waterfall, which made
// the single issuing thread -- not the tensor pipe -- the bottleneck (~700 cycles per 64-wide K block).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_alloc(uint32_t smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Column sums across the 32 lanes of a warp for 32 per-lane values: after the call lane L holds
// sum over lanes of v[L] (in v[0]).  31 shuffles instead of 160.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      float send = up ? v[j] : v[j + s];
      float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

struct SmemLayout {
  uint32_t base;        // 1024-aligned start of the stage ring
  uint32_t full, empty, tfull, tempty, tmem_slot, stat;
};

__device__ __forceinline__ SmemLayout carve(uint8_t* raw, int stages, int stage_bytes) {
  SmemLayout L;
  uint32_t b = (smem_u32(raw) + 1023u) & ~1023u;
  L.base = b;
  uint32_t aux = b + (uint32_t)stages * (uint32_t)stage_bytes;
  L.full = aux;
  L.empty = L.full + 8 * kMaxStages;
  L.tfull = L.empty + 8 * kMaxStages;
  L.tempty = L.tfull + 16;
  L.tmem_slot = L.tempty + 16;
  L.stat = aux + kAuxBytes;           // [4][2][stat_c] floats: per-CTA sum / sum of squares of the outputs, one copy per lane quarter
  return L;
}

// ------------------------------------------------------------------------------------------ kernel A
// Epilogue shared by the conv kernels (warps 2..9): TMEM -> registers -> scale/bias/activation ->
// bf16|fp32 global store, plus the per-CTA BatchNorm statistics.
// pair_rank < 0: one CTA per tile (tile = blockIdx.x, += gridDim.x).  pair_rank in {0,1}: CTA pairs
// (cta_group::2): work item pt = blockIdx.x/2 (+= gridDim.x/2) is two M tiles x one N tile, this CTA owns
// M tile 2*(pt/n_tiles)+rank, and the accumulator-free signal goes to the leader's barrier (cluster address).
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, uint32_t bar_tfull, uint32_t bar_tempty, uint32_t tmem_base,
                                              float* slab, int warp, int lane, int total_tiles, int pair_rank = -1) {
  struct { uint32_t tfull, tempty; } L = {bar_tfull, bar_tempty};
  const int t_first = pair_rank < 0 ? (int)blockIdx.x : (int)(blockIdx.x >> 1);
  const int t_step = pair_rank < 0 ? (int)gridDim.x : (int)(gridDim.x >> 1);
  // ===================================================================== epilogue (warps 2..9)
  // Two warps per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31): they take alternate
  // 32-column chunks, which hides the tcgen05.ld / shuffle latencies of this instruction-heavy stage.
  const int q = warp & 3;                 // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;       // 0: even chunks, 1: odd chunks
  const int row = q * 32 + lane;          // row of the 128-pixel tile
  const int bw = 1 << p.log_bw, bh = 1 << p.log_bh;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.scale) | reinterpret_cast<uintptr_t>(p.bias)) & 15u) == 0;
  int it = 0;
  for (int tile = t_first; tile < total_tiles; tile += t_step, ++it) {
    const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
    const int nt = tile % p.n_tiles;
    int mt = tile / p.n_tiles, phs = 0;
    if (p.phases > 1) { const int per = pair_rank < 0 ? p.m_tiles : p.pair_m; phs = mt / per; mt -= phs * per; }
    if (pair_rank >= 0) mt = 2 * mt + pair_rank;          // an odd tile count leaves a phantom tile: n >= N below
    const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
    const int w = (tw << p.log_bw) + (row & (bw - 1));
    const int h = (th << p.log_bh) + ((row >> p.log_bw) & (bh - 1));
    const int n = (tb << p.log_bn) + (row >> (p.log_bw + p.log_bh));
    const bool valid = (w < p.W) && (h < p.H) && (n < p.N);
    const int n_base = nt * p.block_n;
    int co0 = n_base;
    long long off = (long long)n * p.sN + (long long)h * p.sH + (long long)w * p.sW;
    if (p.phases > 1) off += (long long)(phs >> 1) * p.out_ph + (long long)(phs & 1) * p.out_pw;
    if (p.convt_k) {
      const int tap = n_base / p.Cout;
      co0 = n_base - tap * p.Cout;
      off += (long long)(tap / p.convt_k) * p.sKy + (long long)(tap % p.convt_k) * p.sKx;
    }
    mbar_wait(L.tfull + 8 * as, aph);
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
    for (int cc = half * 32; cc < ((p.ablate & 4) ? 0 : p.block_n); cc += 64) {
      float v[32];
      tc_ld32(t_row + (uint32_t)cc, v);
      const int cbase = co0 + cc;         // first output channel of this chunk
      // whole chunk inside the tile and the tensor (the common case): no per-element bounds selects
      const bool full = (cbase + 32 <= p.Cout) && (cc + 32 <= p.block_n);
      if (p.scale != nullptr && full && vec_ok) {
        // folded eval-mode BatchNorm + activation (the sampling loop's conv): coefficients as 16-byte loads, the affine
        // map and the GELU on packed fp32 pairs.  The scalar form below (64 broadcast loads + ~17 instructions of erf per
        // element) made this epilogue longer than the 27-K-step mainloop of the 192-channel layers: 1.18 instead of 1.62 PFLOP/s.
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + cbase + j));
          const float4 b4 = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + cbase + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          dm::f32x2 A = dm::fma2(dm::pk2(v[j], v[j + 1]), dm::pk2(s4.x, s4.y), dm::pk2(b4.x, b4.y));
          dm::f32x2 B = dm::fma2(dm::pk2(v[j + 2], v[j + 3]), dm::pk2(s4.z, s4.w), dm::pk2(b4.z, b4.w));
          if (p.act == 1) { A = dm::gelu2(A); B = dm::gelu2(B); }
          dm::upk2(A, v[j], v[j + 1]);
          dm::upk2(B, v[j + 2], v[j + 3]);
          if (p.act == 2) { v[j] = fmaxf(v[j], 0.f); v[j + 1] = fmaxf(v[j + 1], 0.f); v[j + 2] = fmaxf(v[j + 2], 0.f); v[j + 3] = fmaxf(v[j + 3], 0.f); }
        }
      } else if (p.scale != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = min(cbase + j, p.Cout - 1);
          v[j] = fmaf(v[j], __ldg(p.scale + c), p.bias != nullptr ? __ldg(p.bias + c) : 0.0f);
        }
      } else if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + min(cbase + j, p.Cout - 1));
      }
      if (!full) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (!(cbase + j < p.Cout && cc + j < p.block_n)) v[j] = 0.0f;
      }
      if (p.scale != nullptr && full && vec_ok) {
        // activation already applied above
      } else if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = dm::gelu_f(v[j]);
      } else if (p.act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
      }
      if (valid) {
        if (p.out_f32) {
          float* o = reinterpret_cast<float*>(p.out) + off + cbase;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (cbase + j < p.ldc_pad && cc + j < p.block_n)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          bf16* o = reinterpret_cast<bf16*>(p.out) + off + cbase;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            if (cbase + j < p.ldc_pad && cc + j < p.block_n) {
              uint4 u;
              u.x = dm::pack2(v[j], v[j + 1]); u.y = dm::pack2(v[j + 2], v[j + 3]);
              u.z = dm::pack2(v[j + 4], v[j + 5]); u.w = dm::pack2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(o + j) = u;
            }
        }
      }
      if (p.stats != nullptr) {
        // per-channel sum and sum of squares over this warp's 32 rows -> combined over the 4 warps below
        // ... of the values as stored (bf16-rounded): what the normalisation pass will read back
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float x0 = v[j], x1 = v[j + 1];
          if (!p.out_f32) { const uint32_t u = dm::pack2(x0, x1); x0 = __uint_as_float(u << 16); x1 = __uint_as_float(u & 0xffff0000u); }
          if (!valid) { x0 = 0.0f; x1 = 0.0f; }
          s1[j] = x0; s2[j] = x0 * x0; s1[j + 1] = x1; s2[j + 1] = x1 * x1;
        }
        const float a = warp_colsum32(s1, lane), b2 = warp_colsum32(s2, lane);
        // accumulated over all of this CTA's tiles in this lane quarter's own copy of the slab: a column of a copy has
        // exactly one writer (warp (q, half) owns the chunks of its parity), so plain adds in tile order -- deterministic
        float* mine = slab + q * 2 * p.stat_c + n_base + cc + lane;
        mine[0] += a;
        mine[p.stat_c] += b2;
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { if (pair_rank < 0) mbar_arrive(L.tempty + 8 * as); else mbar_arrive_cluster(L.tempty + 8 * as); }
  }
  if (p.stats != nullptr) {
    // one partial row per CTA: [blockIdx.x][2][stats_ld]
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int e = threadIdx.x - 64;     // 0..255
    float* g = p.stats + (long long)blockIdx.x * 2 * p.stats_ld;
    const int S = p.stat_c;
    for (int c = e; c < p.Cout; c += 256) {               // the four quarters' copies in a fixed order
      g[c] = (slab[c] + slab[2 * S + c]) + (slab[4 * S + c] + slab[6 * S + c]);
      g[p.stats_ld + c] = (slab[S + c] + slab[3 * S + c]) + (slab[5 * S + c] + slab[7 * S + c]);
    }
    for (int r = blockIdx.x + gridDim.x; r < p.stat_rows; r += gridDim.x) {     // rows no CTA owns: zero
      float* z = p.stats + (long long)r * 2 * p.stats_ld;
      for (int c = e; c < p.Cout; c += 256) { z[c] = 0.0f; z[p.stats_ld + c] = 0.0f; }
    }
  }
}

__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = kAStage + p.b_stage_bytes;
  const SmemLayout L = carve(smem_raw, p.stages, stage_bytes);
  const int num_kb = p.num_taps * (p.chunks0 + p.chunks1);
  const int total_tiles = p.m_tiles * p.n_tiles * (p.phases > 1 ? p.phases : 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(L.full + 8 * s, 1); mbar_init(L.empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(L.tfull + 8 * s, 1); mbar_init(L.tempty + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
  }
  float* const slab = reinterpret_cast<float*>(smem_raw + (L.stat - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 8 * p.stat_c; i += kConvThreads) slab[i] = 0.0f;
  if (warp == 1) tc_alloc(L.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(L.tmem_slot));

  if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform loop)
    {
      int stage = 0; uint32_t phase = 0;
      const int chunks = p.chunks0 + p.chunks1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        int mt = tile / p.n_tiles, phs = 0;
        if (p.phases > 1) { phs = mt / p.m_tiles; mt -= phs * p.m_tiles; }
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw << p.log_bw, h0 = th << p.log_bh, n0 = tb << p.log_bn;
        const int n_base = phs * p.phase_rows + nt * p.block_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(L.empty + 8 * stage, phase ^ 1);
          const uint32_t sa = L.base + stage * stage_bytes, sb = sa + kAStage;
          const uint32_t fb = L.full + 8 * stage;
          const int tap = kb / chunks, ch = kb - tap * chunks;
          const TapInfo t = p.taps[phs * p.num_taps + tap];
          int map = t.map, c0 = ch * kBlockK;
          if (p.dual && ch >= p.chunks0) { map = 1; c0 = (ch - p.chunks0) * kBlockK; }
          if (elect_one()) {
            mbar_expect_tx(fb, kAStage + p.b_stage_bytes);
            tma_load_4d(sa, &p.tmA[map], c0, w0 + t.dx, h0 + t.dy, n0, fb);
            tma_load_2d(sb, &p.tmB, kb * kBlockK, n_base, fb);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (warp-uniform loop)
    {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(L.tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(L.full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = L.base + stage * stage_bytes, sb = sa + kAStage;
          const uint64_t adesc = p.desc_hi | (uint64_t)((sa & 0x3FFFFu) >> 4);
          const uint64_t bdesc = p.desc_hi | (uint64_t)((sb & 0x3FFFFu) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
            tc_commit(L.empty + 8 * stage);
            if (kb == num_kb - 1) tc_commit(L.tfull + 8 * as);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    conv_epilogue(p, L.tfull, L.tempty, tmem_base, slab, warp, lane, total_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tc_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ kernel A2
// CTA-pair version of kernel A (tcgen05.mma.cta_group::2): two M tiles share one weight tile, each CTA stages its own
// 128-pixel activation box and HALF of the weight rows per K step.  Kernel A moves 16 KB + block_n x 128 B per
// 128 x block_n x 64 MACs, which at block_n = 256 needs more L2 -> SM bandwidth than the fabric has (the 4x4 stride-2
// convs and their data gradients ran at 0.6-0.9 PFLOP/s); the pair halves the weight share.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_gemm2_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int b_half = p.b_stage_bytes >> 1;
  const int stage_bytes = kAStage + b_half;
  const SmemLayout L = carve(smem_raw, p.stages, stage_bytes);
  const int chunks = p.chunks0 + p.chunks1;
  const int num_kb = p.num_taps * chunks;
  const int total = p.pair_tiles;
  const int t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(L.full + 8 * s, 1); mbar_init(L.empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(L.tfull + 8 * s, 1); mbar_init(L.tempty + 8 * s, 16); }   // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB2); }
  float* const slab = reinterpret_cast<float*>(smem_raw + (L.stat - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 8 * p.stat_c; i += kConvThreads) slab[i] = 0.0f;
  if (warp == 1) tc_alloc2(L.tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(L.tmem_slot));
  const uint32_t l_full = mapa_shared(L.full, 0), l_tempty = mapa_shared(L.tempty, 0);

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int pt = t_first; pt < total; pt += t_step) {
      const int nt = pt % p.n_tiles;
      int pm = pt / p.n_tiles, phs = 0;
      if (p.phases > 1) { phs = pm / p.pair_m; pm -= phs * p.pair_m; }
      const int mt = 2 * pm + (int)rank;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
      const int w0 = tw << p.log_bw, h0 = th << p.log_bh, n0 = tb << p.log_bn;     // phantom tile: n0 >= N, zero-filled
      const int n_row = phs * p.phase_rows + nt * p.block_n + (int)rank * (p.block_n >> 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(L.empty + 8 * stage, phase ^ 1);
        const uint32_t sa = L.base + stage * stage_bytes, sb = sa + kAStage;
        const int tap = kb / chunks, ch = kb - tap * chunks;
        const TapInfo t = p.taps[phs * p.num_taps + tap];
        int map = t.map, c0 = ch * kBlockK;
        if (p.dual && ch >= p.chunks0) { map = 1; c0 = (ch - p.chunks0) * kBlockK; }
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(L.full + 8 * stage, 2 * stage_bytes);
          tma_load_4d_2sm(sa, &p.tmA[map], c0, w0 + t.dx, h0 + t.dy, n0, l_full + 8 * stage);
          tma_load_2d_2sm(sb, &p.tmB2, kb * kBlockK, n_row, l_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int pt = t_first; pt < total; pt += t_step, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(L.tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(L.full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = L.base + stage * stage_bytes, sb = sa + kAStage;
          const uint64_t adesc = p.desc_hi | (uint64_t)((sa & 0x3FFFFu) >> 4);
          const uint64_t bdesc = p.desc_hi | (uint64_t)((sb & 0x3FFFFu) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              tc_mma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (kb | k) != 0);
            tc_commit2(L.empty + 8 * stage);
            if (kb == num_kb - 1) tc_commit2(L.tfull + 8 * as);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    conv_epilogue(p, L.tfull, l_tempty, tmem_base, slab, warp, lane, total, (int)rank);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tc_dealloc2(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ kernel D
// 3x3 / stride 1 / pad 1 convolution (forward and data gradient) with input-patch reuse: per tile and
// 64-channel chunk ONE TMA box brings the 10 x 18 pixel halo patch into shared memory (SWIZZLE_128B,
// one 128-byte row per pixel); the nine taps are issued as MMAs on nine shifted descriptor views of that
// patch, so the activation operand crosses L2->SM once instead of nine times.  Weight tiles stream
// through their own ring, one [block_n x 64] box per (tap, chunk).
__global__ void __launch_bounds__(kConvThreads, 1) conv3x3_halo_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ringA = base;
  const uint32_t ringB = ringA + (uint32_t)p.a_stages * kHaloStage;
  const uint32_t bars = ringB + (uint32_t)p.stages * (uint32_t)p.b_stage_bytes;
  const uint32_t fullB = bars, emptyB = fullB + 8 * kMaxStages, fullA = emptyB + 8 * kMaxStages, emptyA = fullA + 32,
                 tfull = emptyA + 32, tempty = tfull + 16, tmem_slot = tempty + 16;
  float* const slab = reinterpret_cast<float*>(smem_raw + (bars + kAuxBytes - smem_u32(smem_raw)));
  const int chunks = p.chunks0 + p.chunks1;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 1); }
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(fullA + 8 * s, 1); mbar_init(emptyA + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + 8 * s, 1); mbar_init(tempty + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB); }
  for (int i = threadIdx.x; i < 8 * p.stat_c; i += kConvThreads) slab[i] = 0.0f;
  if (warp == 1) tc_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    {
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw << p.log_bw, h0 = th << p.log_bh;
        const int n_base = nt * p.block_n;
        for (int ch = 0; ch < chunks; ++ch) {
          int map = 0, c0 = ch * kBlockK;
          if (p.dual && ch >= p.chunks0) { map = 1; c0 = (ch - p.chunks0) * kBlockK; }
          mbar_wait(emptyA + 8 * sa, pha ^ 1);
          if (elect_one()) {
            if (p.ablate & 2) mbar_arrive(fullA + 8 * sa);
            else {
              mbar_expect_tx(fullA + 8 * sa, kHaloBytes);
              tma_load_4d(ringA + sa * kHaloStage, &p.tmA[map], c0, w0 - 1, h0 - 1, tb, fullA + 8 * sa);
            }
          }
          __syncwarp();
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(emptyB + 8 * sb, phb ^ 1);
            if (elect_one()) {
              if (p.ablate & 2) mbar_arrive(fullB + 8 * sb);
              else {
                mbar_expect_tx(fullB + 8 * sb, p.b_stage_bytes);
                tma_load_2d(ringB + sb * p.b_stage_bytes, &p.tmB, (tap * chunks + ch) * kBlockK, n_base, fullB + 8 * sb);
              }
            }
            __syncwarp();
            if (++sb == p.stages) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(fullA + 8 * sa, pha);
          const uint32_t a0 = ringA + sa * kHaloStage;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(fullB + 8 * sb, phb);
            tc_fence_after();
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint32_t av = a0 + (uint32_t)((dy * kHaloW + dx) * 128);
            const uint32_t bv = ringB + sb * p.b_stage_bytes;
            uint64_t adesc = p.desc_hi_halo | (uint64_t)((av & 0x3FFFFu) >> 4);
            if (p.base_off_mode == 1) adesc |= (uint64_t)((av >> 7) & 7u) << 49;
            const uint64_t bdesc = p.desc_hi | (uint64_t)((bv & 0x3FFFFu) >> 4);
            const bool last = (ch == chunks - 1) && (tap == 8);
            if (elect_one()) {
              if (p.ablate & 1) {
                mbar_arrive(emptyB + 8 * sb);
                if (tap == 8) mbar_arrive(emptyA + 8 * sa);
                if (last) mbar_arrive(tfull + 8 * as);
              } else {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  tc_mma_bf16((p.ablate & 8) ? tmem_base + (uint32_t)((k & 1) * 256) : d_tmem, adesc + (uint64_t)(k * 2),
                              bdesc + (uint64_t)(k * 2), p.idesc, (ch | tap | k) != 0);   // ablate bit3: two independent chains
                tc_commit(emptyB + 8 * sb);
                if (tap == 8) tc_commit(emptyA + 8 * sa);
                if (last) tc_commit(tfull + 8 * as);
              }
            }
            __syncwarp();
            if (++sb == p.stages) { sb = 0; phb ^= 1; }
          }
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else {
    conv_epilogue(p, tfull, tempty, tmem_base, slab, warp, lane, total_tiles);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tc_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ kernel E
// CTA-pair version of kernel D (tcgen05.mma.cta_group::2, M = 256 per instruction).  A single-CTA 128 x N MMA
// stream is capped by the MMA issue rate (~167 cycles per instruction whatever N: 1.48 PFLOP/s at N=192 with
// loads and epilogue switched off); a pair instruction does twice the work, and each CTA stages only HALF of
// the weight tile.  Two CTAs of a cluster take two M tiles with the same N tile: each loads its own halo
// patch and block_n/2 weight rows (TMA with the LEADER's mbarrier as completion target), the leader's MMA
// warp issues for both, tcgen05.commit multicasts the stage-free / accumulator-ready arrivals to both CTAs,
// and both CTAs run the epilogue on their own 128 TMEM lanes.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv3x3_halo2_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int b_half = p.b_stage_bytes >> 1;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ringA = base;
  const uint32_t ringB = ringA + (uint32_t)p.a_stages * kHaloStage;
  const uint32_t bars = ringB + (uint32_t)p.stages * (uint32_t)b_half;
  const uint32_t fullB = bars, emptyB = fullB + 8 * kMaxStages, fullA = emptyB + 8 * kMaxStages, emptyA = fullA + 32,
                 tfull = emptyA + 32, tempty = tfull + 16, tmem_slot = tempty + 16;
  float* const slab = reinterpret_cast<float*>(smem_raw + (bars + kAuxBytes - smem_u32(smem_raw)));
  const int chunks = p.chunks0 + p.chunks1;
  const int total = p.pair_tiles;
  const int t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 1); }
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(fullA + 8 * s, 1); mbar_init(emptyA + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + 8 * s, 1); mbar_init(tempty + 8 * s, 16); }   // 8 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB2); }
  for (int i = threadIdx.x; i < 8 * p.stat_c; i += kConvThreads) slab[i] = 0.0f;
  if (warp == 1) tc_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                      // both CTAs' barriers initialised and TMEM allocated
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // the leader's (rank 0) barriers as shared::cluster addresses
  const uint32_t l_fullA = mapa_shared(fullA, 0), l_fullB = mapa_shared(fullB, 0), l_tempty = mapa_shared(tempty, 0);

  if (warp == 0) {
    int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
    for (int pt = t_first; pt < total; pt += t_step) {
      const int nt = pt % p.n_tiles, mt = 2 * (pt / p.n_tiles) + (int)rank;
      const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
      const int w0 = tw << p.log_bw, h0 = th << p.log_bh;
      const int n_row = nt * p.block_n + (int)rank * (p.block_n >> 1);
      for (int ch = 0; ch < chunks; ++ch) {
        int map = 0, c0 = ch * kBlockK;
        if (p.dual && ch >= p.chunks0) { map = 1; c0 = (ch - p.chunks0) * kBlockK; }
        mbar_wait(emptyA + 8 * sa, pha ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(fullA + 8 * sa, 2 * kHaloBytes);
          tma_load_4d_2sm(ringA + sa * kHaloStage, &p.tmA[map], c0, w0 - 1, h0 - 1, tb, l_fullA + 8 * sa);
        }
        __syncwarp();
        if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(emptyB + 8 * sb, phb ^ 1);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(fullB + 8 * sb, 2 * b_half);
            tma_load_2d_2sm(ringB + sb * b_half, &p.tmB2, (tap * chunks + ch) * kBlockK, n_row, l_fullB + 8 * sb);
          }
          __syncwarp();
          if (++sb == p.stages) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int pt = t_first; pt < total; pt += t_step, ++it) {
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int ch = 0; ch < chunks; ++ch) {
          mbar_wait(fullA + 8 * sa, pha);
          const uint32_t a0 = ringA + sa * kHaloStage;
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(fullB + 8 * sb, phb);
            tc_fence_after();
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint32_t av = a0 + (uint32_t)((dy * kHaloW + dx) * 128);
            const uint32_t bv = ringB + sb * b_half;
            const uint64_t adesc = p.desc_hi_halo | (uint64_t)((av & 0x3FFFFu) >> 4);
            const uint64_t bdesc = p.desc_hi | (uint64_t)((bv & 0x3FFFFu) >> 4);
            const bool last = (ch == chunks - 1) && (tap == 8);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                tc_mma2_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc, (ch | tap | k) != 0);
              tc_commit2(emptyB + 8 * sb);
              if (tap == 8) tc_commit2(emptyA + 8 * sa);
              if (last) tc_commit2(tfull + 8 * as);
            }
            __syncwarp();
            if (++sb == p.stages) { sb = 0; phb ^= 1; }
          }
          if (++sa == p.a_stages) { sa = 0; pha ^= 1; }
        }
      }
    }
  } else {
    conv_epilogue(p, tfull, l_tempty, tmem_base, slab, warp, lane, total, (int)rank);
  }
  tc_fence_before();
  cluster_sync_all();                      // nobody leaves while the peer may still read its smem / signal its barriers
  if (warp == 1) { tc_fence_after(); tc_dealloc2(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ kernel B
__global__ void __launch_bounds__(kThreads, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = kAStage + p.b_stage_bytes;   // A: 2 boxes of [64 px][64 co]; B: block_n/64 boxes
  const SmemLayout L = carve(smem_raw, p.stages, stage_bytes);
  const int chunks = p.chunks0 + p.chunks1;            // 64-channel chunks of Cin
  const int cpt = p.block_n / 64;                      // chunks per ci tile
  const int total_tiles = p.co_tiles * p.num_taps * p.ci_tiles * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(L.full + 8 * s, 1); mbar_init(L.empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(L.tfull + 8 * s, 1); mbar_init(L.tempty + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tc_alloc(L.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(L.tmem_slot));

  // tile -> (split, ci tile, tap, co tile); co fastest so concurrent CTAs share the X patches in L2
  auto decode = [&](int tile, int& cot, int& tap, int& cit, int& sp) {
    cot = tile % p.co_tiles; tile /= p.co_tiles;
    tap = tile % p.num_taps; tile /= p.num_taps;
    cit = tile % p.ci_tiles; sp = tile / p.ci_tiles;
  };

  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int cot, tap, cit, sp; decode(tile, cot, tap, cit, sp);
        const TapInfo t = p.taps[tap];
        const int p0 = sp * p.patches_per_split;
        const int p1 = min(p.patches, p0 + p.patches_per_split);
        const int nch = min(cpt, chunks - cit * cpt);   // live 64-channel boxes of this ci tile
        for (int pp = p0; pp < p1; ++pp) {
          const int tw = pp % p.tiles_w, th = (pp / p.tiles_w) % p.tiles_h, tb = pp / (p.tiles_w * p.tiles_h);
          const int w0 = tw << p.log_bw, h0 = th << p.log_bh, n0 = tb << p.log_bn;
          mbar_wait(L.empty + 8 * stage, phase ^ 1);
          const uint32_t sa = L.base + stage * stage_bytes, sb = sa + kAStage;
          const uint32_t fb = L.full + 8 * stage;
          if (elect_one()) {
            mbar_expect_tx(fb, 2 * 8192 + nch * 8192);
            tma_load_4d(sa, &p.tmDY, cot * 128, w0, h0, n0, fb);
            tma_load_4d(sa + 8192, &p.tmDY, cot * 128 + 64, w0, h0, n0, fb);
            for (int j = 0; j < nch; ++j) {
              const int ch = cit * cpt + j;
              int map = t.map, c0 = ch * 64;
              if (p.dual && ch >= p.chunks0) { map = 1; c0 = (ch - p.chunks0) * 64; }
              tma_load_4d(sb + j * 8192, &p.tmX[map], c0, w0 + t.dx, h0 + t.dy, n0, fb);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int cot, tap, cit, sp; decode(tile, cot, tap, cit, sp);
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        const int p0 = sp * p.patches_per_split;
        const int p1 = min(p.patches, p0 + p.patches_per_split);
        mbar_wait(L.tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int pp = p0; pp < p1; ++pp) {
          mbar_wait(L.full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = L.base + stage * stage_bytes, sb = sa + kAStage;
          const uint64_t adesc = p.desc_hi_a | ((uint64_t)p.lbo_a << 16) | (uint64_t)((sa & 0x3FFFFu) >> 4);
          const uint64_t bdesc = p.desc_hi_b | ((uint64_t)p.lbo_b << 16) | (uint64_t)((sb & 0x3FFFFu) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)      // 16 pixels per MMA = two 8-row swizzle groups = 2048 B
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), p.idesc, (pp != p0) | (k != 0));
            tc_commit(L.empty + 8 * stage);
            if (pp == p1 - 1) tc_commit(L.tfull + 8 * as);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int cot, tap, cit, sp; decode(tile, cot, tap, cit, sp);
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      const int co = cot * 128 + q * 32 + lane;
      const int ci0 = cit * p.block_n;
      const int ncols = min(p.block_n, p.cin_k - ci0);
      mbar_wait(L.tfull + 8 * as, aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      float* dst = p.dwp + ((long long)co * p.num_taps + tap) * p.cin_k + ci0;
      for (int cc = 0; cc < ncols; cc += 32) {
        float v[32];
        tc_ld32(t_row + (uint32_t)cc, v);
        if (co < p.Cout) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (cc + j < ncols)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + cc + j), "f"(v[j]), "f"(v[j + 1]),
                           "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(L.tempty + 8 * as);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tc_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ kernel C
__global__ void __launch_bounds__(kThreads, 1) wgrad2_gemm_kernel(const __grid_constant__ Wgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = 4 * 8192, b_bytes = p.nb * 8192;
  const int stage_bytes = a_bytes + b_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + (uint32_t)p.stages * (uint32_t)stage_bytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxStages, bar_tfull = bar_empty + 8 * kMaxStages,
                 bar_tempty = bar_tfull + 8, tmem_slot = bar_tempty + 8;
  const int chunks = p.chunks0 + p.chunks1;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tfull, 1); mbar_init(bar_tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmX[0]); tma_prefetch_desc(&p.tmDY); }
  if (warp == 1) tc_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // tile -> (split, n tile, m tile); m fastest: CTAs running together sweep the same pixel range, so
  // the dY tile and the (overlapping) shifted X boxes are shared through L2
  auto decode = [&](int tile, int& mt, int& nt, int& sp) {
    mt = tile % p.m_tiles; tile /= p.m_tiles;
    nt = tile % p.n_tiles; sp = tile / p.n_tiles;
  };

  if (warp == 0) {
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt, nt, sp; decode(tile, mt, nt, sp);
        const int nA = min(4, p.total_boxes - mt * 4);
        const int p0 = sp * p.patches_per_split;
        const int p1 = min(p.patches, p0 + p.patches_per_split);
        int bmap[4], bc0[4], bdx[4], bdy[4];
        for (int j = 0; j < 4; ++j) {
          const int box = min(mt * 4 + j, p.total_boxes - 1);
          const int tap = box / chunks, ch = box - tap * chunks;
          const TapInfo t = p.taps[tap];
          bmap[j] = t.map; bc0[j] = ch * 64; bdx[j] = t.dx; bdy[j] = t.dy;
          if (p.dual && ch >= p.chunks0) { bmap[j] = 1; bc0[j] = (ch - p.chunks0) * 64; }
        }
        for (int pp = p0; pp < p1; ++pp) {
          const int tw = pp % p.tiles_w, th = (pp / p.tiles_w) % p.tiles_h, tb = pp / (p.tiles_w * p.tiles_h);
          const int w0 = tw << p.log_bw, h0 = th << p.log_bh, n0 = tb << p.log_bn;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
          const uint32_t fb = bar_full + 8 * stage;
          if (elect_one()) {
            mbar_expect_tx(fb, (nA + p.nb) * 8192);
            for (int j = 0; j < p.nb; ++j) tma_load_4d(sb + j * 8192, &p.tmDY, nt * p.block_n + j * 64, w0, h0, n0, fb);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < nA) tma_load_4d(sa + j * 8192, &p.tmX[bmap[j]], bc0[j], w0 + bdx[j], h0 + bdy[j], n0, fb);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        int mt, nt, sp; decode(tile, mt, nt, sp);
        const int nA = min(4, p.total_boxes - mt * 4);
        const int nh = nA > 2 ? 2 : 1;
        const int p0 = sp * p.patches_per_split;
        const int p1 = min(p.patches, p0 + p.patches_per_split);
        mbar_wait(bar_tempty, (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        for (int pp = p0; pp < p1; ++pp) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
          const uint64_t bdesc = p.desc_hi | ((uint64_t)p.lbo << 16) | (uint64_t)((sb & 0x3FFFFu) >> 4);
          const uint64_t adesc0 = p.desc_hi | ((uint64_t)p.lbo << 16) | (uint64_t)((sa & 0x3FFFFu) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {    // 16 pixels per MMA = two 8-row swizzle groups = 2048 B
              tc_mma_bf16(tmem_base, adesc0 + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), p.idesc, (pp != p0) | (k != 0));
              if (nh == 2)                   // second accumulator: boxes 2,3 (16 KB further = 1024 descriptor units)
                tc_mma_bf16(tmem_base + 256u, adesc0 + (uint64_t)(1024 + k * 128), bdesc + (uint64_t)(k * 128), p.idesc,
                            (pp != p0) | (k != 0));
            }
            tc_commit(bar_empty + 8 * stage);
            if (pp == p1 - 1) tc_commit(bar_tfull);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int mt, nt, sp; decode(tile, mt, nt, sp);
      const int nA = min(4, p.total_boxes - mt * 4);
      const int nh = nA > 2 ? 2 : 1;
      mbar_wait(bar_tfull, (uint32_t)(it & 1));
      tc_fence_after();
      for (int h = 0; h < nh; ++h) {
        const int box = mt * 4 + 2 * h + (r >> 6);
        const int tap = box / chunks, ch = box - tap * chunks;
        const int cil = r & 63;
        const bool first = !(p.dual && ch >= p.chunks0);
        const int creal = first ? ch * 64 + cil : (ch - p.chunks0) * 64 + cil;
        const bool live = box < p.total_boxes && creal < (first ? p.C0 : p.C1);
        float* dst = p.dwp + (long long)tap * p.cin_k + ch * 64 + cil;
        const long long co_stride = (long long)p.num_taps * p.cin_k;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 256);
        for (int cc = 0; cc < p.block_n; cc += 32) {
          float v[32];
          tc_ld32(t_row + (uint32_t)cc, v);
          if (live) {
            const int co0 = nt * p.block_n + cc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (co0 + j < p.Cout)
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long long)(co0 + j) * co_stride), "f"(v[j]) : "memory");
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tc_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ kernel F
// CTA-pair weight gradient (cta_group::2): the quad of im2col boxes of an M tile is split over the pair
// (rank r stages boxes 2r, 2r+1 = its 128 accumulator rows) and each CTA stages half of the dY tile's
// channels; one M=256 instruction per 16 pixels does the work of kernel C's two.  A CTA now holds ONE
// 128 x N accumulator, so TMEM is double-buffered again and the red.add epilogue overlaps the next tile.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
wgrad3_pair_kernel(const __grid_constant__ Wgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int nbh = p.nb;                                  // dY boxes per CTA (its half of the tile's channels)
  const int a_bytes = 2 * 8192, stage_bytes = a_bytes + nbh * 8192;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + (uint32_t)p.stages * (uint32_t)stage_bytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxStages, bar_tfull = bar_empty + 8 * kMaxStages,
                 bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  const int chunks = p.chunks0 + p.chunks1;
  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;
  const int t_first = blockIdx.x >> 1, t_step = gridDim.x >> 1;
  const int half_n = p.block_n >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 8); }   // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmX[0]); tma_prefetch_desc(&p.tmDY); }
  if (warp == 1) tc_alloc2(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t l_full = mapa_shared(bar_full, 0), l_tempty = mapa_shared(bar_tempty, 0);

  auto decode = [&](int tile, int& mt, int& nt, int& sp) {
    mt = tile % p.m_tiles; tile /= p.m_tiles;
    nt = tile % p.n_tiles; sp = tile / p.n_tiles;
  };

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = t_first; tile < total_tiles; tile += t_step) {
      int mt, nt, sp; decode(tile, mt, nt, sp);
      const int nA_all = min(4, p.total_boxes - mt * 4);              // live boxes of the quad
      const int nA = max(0, min(2, nA_all - 2 * (int)rank));          // ... of which this CTA stages
      const int p0 = sp * p.patches_per_split;
      const int p1 = min(p.patches, p0 + p.patches_per_split);
      int bmap[2], bc0[2], bdx[2], bdy[2];
      for (int j = 0; j < 2; ++j) {
        const int box = min(mt * 4 + 2 * (int)rank + j, p.total_boxes - 1);
        const int tap = box / chunks, ch = box - tap * chunks;
        const TapInfo t = p.taps[tap];
        bmap[j] = t.map; bc0[j] = ch * 64; bdx[j] = t.dx; bdy[j] = t.dy;
        if (p.dual && ch >= p.chunks0) { bmap[j] = 1; bc0[j] = (ch - p.chunks0) * 64; }
      }
      const int co0 = nt * p.block_n + (int)rank * half_n;
      for (int pp = p0; pp < p1; ++pp) {
        const int tw = pp % p.tiles_w, th = (pp / p.tiles_w) % p.tiles_h, tb = pp / (p.tiles_w * p.tiles_h);
        const int w0 = tw << p.log_bw, h0 = th << p.log_bh, n0 = tb << p.log_bn;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
        const uint32_t fb = l_full + 8 * stage;
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, (nA_all + 2 * nbh) * 8192);
          for (int j = 0; j < nbh; ++j) tma_load_4d_2sm(sb + j * 8192, &p.tmDY, co0 + j * 64, w0, h0, n0, fb);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (j < nA) tma_load_4d_2sm(sa + j * 8192, &p.tmX[bmap[j]], bc0[j], w0 + bdx[j], h0 + bdy[j], n0, fb);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = t_first; tile < total_tiles; tile += t_step, ++it) {
        int mt, nt, sp; decode(tile, mt, nt, sp);
        const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
        const int p0 = sp * p.patches_per_split;
        const int p1 = min(p.patches, p0 + p.patches_per_split);
        mbar_wait(bar_tempty + 8 * as, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
        for (int pp = p0; pp < p1; ++pp) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * stage_bytes, sb = sa + a_bytes;
          const uint64_t adesc = p.desc_hi | ((uint64_t)p.lbo << 16) | (uint64_t)((sa & 0x3FFFFu) >> 4);
          const uint64_t bdesc = p.desc_hi | ((uint64_t)p.lbo << 16) | (uint64_t)((sb & 0x3FFFFu) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)      // 16 pixels per MMA = two 8-row swizzle groups = 2048 B
              tc_mma2_bf16(d_tmem, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), p.idesc, (pp != p0) | (k != 0));
            tc_commit2(bar_empty + 8 * stage);
            if (pp == p1 - 1) tc_commit2(bar_tfull + 8 * as);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = t_first; tile < total_tiles; tile += t_step, ++it) {
      int mt, nt, sp; decode(tile, mt, nt, sp);
      const int as = it & 1; const uint32_t aph = (it >> 1) & 1;
      mbar_wait(bar_tfull + 8 * as, aph);
      tc_fence_after();
      const int box = mt * 4 + 2 * (int)rank + (r >> 6);
      const int tap = box / chunks, ch = box - tap * chunks;
      const int cil = r & 63;
      const bool first = !(p.dual && ch >= p.chunks0);
      const int creal = first ? ch * 64 + cil : (ch - p.chunks0) * 64 + cil;
      const bool live = box < p.total_boxes && creal < (first ? p.C0 : p.C1);
      float* dst = p.dwp + (long long)tap * p.cin_k + ch * 64 + cil;
      const long long co_stride = (long long)p.num_taps * p.cin_k;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
      for (int cc = 0; cc < p.block_n; cc += 32) {
        float v[32];
        tc_ld32(t_row + (uint32_t)cc, v);
        if (live) {
          const int co0 = nt * p.block_n + cc;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (co0 + j < p.Cout && cc + j < p.block_n)
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + (long long)(co0 + j) * co_stride), "f"(v[j]) : "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_tempty + 8 * as);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tc_dealloc2(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int g_sm_limit = DM_NUM_SMS;   // SMs the persistent GEMM grids may occupy (dm_set_sm_limit)
long long g_debug[16] = {0};   // 8..15: elementwise.cu experiments (dm_debug_value); 0 stages 1 grid 2 splits 3 block_n 4 wgrad kernel (1 B, 2 C, 3 F) 5 no-halo 6 base-offset 7 ablation 9 upcat per-pixel 10 wgrad stages

int ensure_encode() {
  if (g_encode) return DM_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr) { dm_set_error("cuTensorMapEncodeTiled entry point not found"); return DM_ERR_TMA; }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return DM_OK;
}

// 4-D map over an NHWC bf16 activation view: dims (C, W, H, N), element strides (1, sw, sh, sn).
int make_act_map(CUtensorMap* m, const void* base, int C, int W, int H, int N, long long sw, long long sh, long long sn,
                 int bw, int bh, int bn) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(act) failed: %d (C=%d W=%d H=%d N=%d sw=%lld sh=%lld sn=%lld box=%d,%d,%d ptr=%p)",
             (int)r, C, W, H, N, sw, sh, sn, bw, bh, bn, base);
    dm_set_error(buf);
    return DM_ERR_TMA;
  }
  return DM_OK;
}

// 2-D map over the packed weight matrix [rows, ktot] (K contiguous).
int make_w_map(CUtensorMap* m, const void* base, long long rows, long long ktot, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(weights) failed: %d (rows=%lld ktot=%lld box=%d)", (int)r, rows, ktot, box_rows);
    dm_set_error(buf);
    return DM_ERR_TMA;
  }
  return DM_OK;
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
int pow2ceil(int v) { return 1 << ilog2(v); }

void pick_patch(int W, int H, int rows, int& lbw, int& lbh, int& lbn) {
  int bw = pow2ceil(W); if (bw > rows) bw = rows;
  int bh = pow2ceil(H); if (bh > rows / bw) bh = rows / bw;
  int bn = rows / (bw * bh);
  lbw = ilog2(bw); lbh = ilog2(bh); lbn = ilog2(bn);
}

int pick_block_n(int cout) {
  if (g_debug[3] >= 16 && g_debug[3] <= 256 && (g_debug[3] % 16) == 0) return (int)g_debug[3];   // dev: forced tile width
  int c16 = (cout + 15) / 16 * 16;
  int nt = (c16 + 255) / 256;
  int bn = ((c16 + nt - 1) / nt + 15) / 16 * 16;
  return bn;
}

// Tile width for a conv with `m_tiles` 128-pixel tiles: the persistent grid runs ceil(tiles / #SMs) rounds of
// (block_n + fixed) cost each, so for the small-spatial layers (32 M-tiles at 32x32, batch 4) a narrower
// tile that fills the last round beats the widest one: Cout=768 -> 4 x 192 instead of 3 x 256.
// pairs: the 3x3 halo convs run as CTA pairs (two M tiles share one weight tile, half of it staged by each CTA) whenever
// the width is a multiple of 32 -- a single CTA staging a whole 176-wide weight tile per K step is bound by the
// L2 -> SM fabric (1.56 GB in 154 us = 10.1 TB/s on 1536 -> 1536 at 32x32, tensor pipe 64 % active), so the pick is made
// among the pair-capable widths with the pair kernel's own round count.
int pick_block_n_tiles(int cout, int m_tiles, bool pairs = false) {
  if (g_debug[3] >= 16 && g_debug[3] <= 256 && (g_debug[3] % 16) == 0) return (int)g_debug[3];
  const int widest = pick_block_n(cout);
  if (cout <= 256) return widest;
  int best = widest; double best_cost = 1e30;
  const int slots = pairs ? g_sm_limit / 2 : g_sm_limit;
  const long long m_items = pairs ? (m_tiles + 1) / 2 : m_tiles;
  for (int bn = 256; bn >= 96; bn -= (pairs ? 32 : 16)) {
    const int nt = (cout + bn - 1) / bn;
    const long long tiles = m_items * nt;
    const long long rounds = (tiles + slots - 1) / slots;
    const double cost = (double)rounds * (bn + 48);      // 48 ~ the per-tile share that does not shrink with N
    if (cost < best_cost * 0.98) { best_cost = cost; best = bn; }
  }
  return best;
}

unsigned make_idesc(int n, bool a_mn, bool b_mn) {
  // kind::f16: c=f32 (1<<4), a=bf16 (1<<7), b=bf16 (1<<10), a_major<<15, b_major<<16, N>>3 <<17, M>>4 <<24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((unsigned)(n >> 3) << 17) | ((unsigned)(kBlockM >> 4) << 24);
}

// SBO = 1024 B (8 rows x 128 B), descriptor version 1 (sm_100), SWIZZLE_128B
constexpr unsigned long long kDescHiKMajor = (64ull << 32) | (1ull << 46) | (2ull << 61) | (1ull << 16);
constexpr unsigned long long kDescHiMN = (64ull << 32) | (1ull << 46) | (2ull << 61);

bool g_attr_a = false, g_attr_b = false;

struct ConvGeom {
  // one launch of kernel A
  const void* src[4]; int srcC[4]; long long src_sw[4], src_sh[4], src_sn[4]; int src_W[4], src_H[4];
  int num_src;
};

}  // namespace

extern "C" int dm_conv2d_fwd_stat_rows(int N, int Ho, int Wo, int Cout);
// Leave SMs free for a concurrent kernel on another stream (an NCCL all-reduce overlapped with the backward pass): the
// GEMM kernels hold a whole SM's shared memory per CTA, so a full-width grid would queue behind such a kernel's CTAs.
extern "C" int dm_set_sm_limit(int sms) {
  if (sms < 2 || sms > DM_NUM_SMS) { dm_set_error("dm_set_sm_limit: 2..148"); return DM_ERR_ARG; }
  g_sm_limit = sms & ~1;          // CTA pairs
  return DM_OK;
}
extern "C" void dm_debug_set(int key, long long value) { if (key >= 0 && key < 16) g_debug[key] = value; }
long long dm_debug_value(int key) { return (key >= 0 && key < 16) ? g_debug[key] : 0; }

// Generic launcher for kernel A.  All geometry is resolved by the typed entry points below.
static int conv_grid(int tiles) { return tiles < g_sm_limit ? tiles : g_sm_limit; }

// CTA pairs pay when the M tiles pair up without (much of) a phantom tile
static bool pair_ok(int m_tiles) { return g_debug[5] != 2 && m_tiles >= 2 && ((m_tiles & 1) == 0 || m_tiles >= 9); }

bool g_attr_a2 = false;

// wpk / w_rows / ktot given: the launch may run as CTA pairs (kernel A2) when the tile width and tile count allow it.
static int launch_conv(ConvParams& P, cudaStream_t st, const void* wpk = nullptr, long long w_rows = 0, long long ktot = 0) {
  const int phases = P.phases > 1 ? P.phases : 1;
  const bool pair = wpk != nullptr && (P.block_n % 32) == 0 && pair_ok(P.m_tiles);
  const int b_bytes = pair ? P.b_stage_bytes / 2 : P.b_stage_bytes;
  int stage_bytes = kAStage + b_bytes;
  P.stat_c = P.stats ? (P.n_tiles * P.block_n + 31) / 32 * 32 : 0;   // whole 32-lane chunks
  const int stat_bytes = 32 * P.stat_c;            // [4 lane quarters][2][stat_c] floats
  if (stat_bytes > kMaxStatBytes) { dm_set_error("conv_gemm: too many output channels for fused BatchNorm statistics"); return DM_ERR_ARG; }
  int stages = (kSmemBudget - 1024 - kAuxBytes - stat_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (g_debug[0] > 0 && g_debug[0] < stages) stages = (int)g_debug[0];
  if (stages < 2) { dm_set_error("conv_gemm: not enough shared memory for 2 stages"); return DM_ERR_ARG; }
  P.stages = stages;
  P.ablate = (int)g_debug[7] & 4;          // generic kernel: only the epilogue switch
  size_t smem = 1024 + (size_t)stages * stage_bytes + kAuxBytes + stat_bytes;
  if (pair) {
    int rc = make_w_map(&P.tmB2, wpk, w_rows, ktot, P.block_n / 2);
    if (rc) return rc;
    P.idesc = (P.idesc & ~(0x1Fu << 24)) | ((256u >> 4) << 24);       // M = 256 across the CTA pair
    P.pair_m = dm::cdiv(P.m_tiles, 2);
    P.pair_tiles = P.pair_m * P.n_tiles * phases;
    P.ablate = 0;
    if (!g_attr_a2) {
      cudaError_t e = cudaFuncSetAttribute(conv_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
      g_attr_a2 = true;
    }
    const int clusters = P.pair_tiles < g_sm_limit / 2 ? P.pair_tiles : g_sm_limit / 2;
    dm_note_kernel("conv_gemm2", P.block_n);
    conv_gemm2_kernel<<<2 * clusters, kConvThreads, smem, st>>>(P);
    DM_CHECK_LAUNCH();
    return DM_OK;
  }
  if (!g_attr_a) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
    g_attr_a = true;
  }
  int grid = conv_grid(P.m_tiles * P.n_tiles * phases);
  if (!P.stats && g_debug[1] > 0 && g_debug[1] < grid) grid = (int)g_debug[1];
  dm_note_kernel("conv_gemm", P.block_n);
  conv_gemm_kernel<<<grid, kConvThreads, smem, st>>>(P);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

bool g_attr_d = false;
// SBO = 1280 B: consecutive 8-pixel rows of the tile are one halo row (10 pixels) apart
constexpr unsigned long long kDescHiHalo = ((unsigned long long)(kHaloW * 128 / 16) << 32) | (1ull << 46) | (2ull << 61) | (1ull << 16);

static int launch_conv_halo(ConvParams& P, cudaStream_t st) {
  P.stat_c = P.stats ? (P.n_tiles * P.block_n + 31) / 32 * 32 : 0;
  const int stat_bytes = 32 * P.stat_c;            // [4 lane quarters][2][stat_c] floats
  if (stat_bytes > kMaxStatBytes) { dm_set_error("conv_gemm: too many output channels for fused BatchNorm statistics"); return DM_ERR_ARG; }
  P.a_stages = kHaloAStages;
  int stages = (kSmemBudget - 1024 - kAuxBytes - stat_bytes - P.a_stages * kHaloStage) / P.b_stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (g_debug[0] > 0 && g_debug[0] < stages) stages = (int)g_debug[0];
  if (stages < 2) { dm_set_error("conv3x3_halo: not enough shared memory"); return DM_ERR_ARG; }
  P.stages = stages;
  P.desc_hi_halo = kDescHiHalo;
  P.base_off_mode = (int)g_debug[6];
  P.ablate = (int)g_debug[7];
  const size_t smem = 1024 + (size_t)P.a_stages * kHaloStage + (size_t)stages * P.b_stage_bytes + kAuxBytes + stat_bytes;
  if (!g_attr_d) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
    g_attr_d = true;
  }
  const int grid = conv_grid(P.m_tiles * P.n_tiles);
  dm_note_kernel("conv3x3_halo", P.block_n);
  conv3x3_halo_kernel<<<grid, kConvThreads, smem, st>>>(P);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

bool g_attr_e = false;
static int launch_conv_halo2(ConvParams& P, const void* wpk, long long w_rows, long long ktot, cudaStream_t st) {
  P.stat_c = P.stats ? (P.n_tiles * P.block_n + 31) / 32 * 32 : 0;
  const int stat_bytes = 32 * P.stat_c;            // [4 lane quarters][2][stat_c] floats
  if (stat_bytes > kMaxStatBytes) { dm_set_error("conv_gemm: too many output channels for fused BatchNorm statistics"); return DM_ERR_ARG; }
  int rc = make_w_map(&P.tmB2, wpk, w_rows, ktot, P.block_n / 2);
  if (rc) return rc;
  P.a_stages = kHaloAStages;
  const int b_half = P.b_stage_bytes / 2;
  int stages = (kSmemBudget - 1024 - kAuxBytes - stat_bytes - P.a_stages * kHaloStage) / b_half;
  if (stages > kMaxStages) stages = kMaxStages;
  if (g_debug[0] > 0 && g_debug[0] < stages) stages = (int)g_debug[0];
  P.stages = stages;
  P.desc_hi_halo = kDescHiHalo;
  P.base_off_mode = 0; P.ablate = 0;
  P.idesc = (P.idesc & ~(0x1Fu << 24)) | ((256u >> 4) << 24);       // M = 256 across the CTA pair
  P.pair_m = dm::cdiv(P.m_tiles, 2);
  P.pair_tiles = P.pair_m * P.n_tiles;
  const size_t smem = 1024 + (size_t)P.a_stages * kHaloStage + (size_t)stages * b_half + kAuxBytes + stat_bytes;
  if (!g_attr_e) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
    g_attr_e = true;
  }
  const int clusters = P.pair_tiles < g_sm_limit / 2 ? P.pair_tiles : g_sm_limit / 2;
  dm_note_kernel("conv3x3_halo2", P.block_n);
  conv3x3_halo2_kernel<<<2 * clusters, kConvThreads, smem, st>>>(P);
  DM_CHECK_LAUNCH();
  return DM_OK;
}

static void fill_common(ConvParams& P, int N, int H, int W, int Cout_rows, int block_n) {
  pick_patch(W, H, kBlockM, P.log_bw, P.log_bh, P.log_bn);
  P.tiles_w = dm::cdiv(W, 1 << P.log_bw);
  P.tiles_h = dm::cdiv(H, 1 << P.log_bh);
  P.tiles_b = dm::cdiv(N, 1 << P.log_bn);
  P.m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  P.block_n = block_n;
  P.n_tiles = dm::cdiv(Cout_rows, block_n);
  P.N = N; P.H = H; P.W = W;
  P.b_stage_bytes = block_n * 128;
  P.idesc = make_idesc(block_n, false, false);
  P.desc_hi = kDescHiKMajor;
}

// ---------------------------------------------------------------------------------------------------
// dm_conv2d_fwd: y[N,Ho,Wo,Cout] = act(conv(x0 (++ x1 on channels), Wp) * scale + bias)   (bf16 NHWC in, bf16/fp32 out)
// scale (nullable -> 1) and act (0 none / 1 GELU / 2 ReLU) fold an eval-mode BatchNorm + activation into the epilogue.
// stride 1: any kh,kw,pad.  stride 2: even Hin,Win (parity views), single source.
// stats (optional): per-CTA partial sums [dm_conv2d_fwd_stat_rows()][2][stats_ld] of y and y^2 for train-mode BatchNorm.
extern "C" int dm_conv2d_fwd(const void* x0, int C0, int ld0, const void* x1, int C1, int ld1, const void* wpk,
                             const float* bias, const float* scale, int act, void* y, int ldy, int y_f32, float* stats,
                             int stats_ld, int N, int Hin, int Win, int Cout, int kh, int kw, int stride, int pad,
                             void* stream) {
  if (ensure_encode() != DM_OK) return DM_ERR_TMA;
  if (kh * kw > 16 || (stride != 1 && stride != 2)) { dm_set_error("dm_conv2d_fwd: unsupported kernel/stride"); return DM_ERR_ARG; }
  if (stride == 2 && (x1 != nullptr || (Hin & 1) || (Win & 1))) { dm_set_error("dm_conv2d_fwd: stride 2 needs one source, even H/W"); return DM_ERR_ARG; }
  if ((ld0 & 7) || (x1 && (ld1 & 7)) || (ldy & (y_f32 ? 3 : 7))) { dm_set_error("dm_conv2d_fwd: channel pitch must be a multiple of 8"); return DM_ERR_ARG; }
  const int Ho = (Hin + 2 * pad - kh) / stride + 1, Wo = (Win + 2 * pad - kw) / stride + 1;
  ConvParams P;
  memset(&P, 0, sizeof P);
  // 3x3 / stride 1 / pad 1 on images that tile into 8 x 16 pixel patches: the halo kernel
  bool halo = kh == 3 && kw == 3 && stride == 1 && pad == 1 && (Win % 8) == 0 && (Hin % 16) == 0;
  if (g_debug[5] == 1) halo = false;
  int lbw, lbh, lbn;
  pick_patch(Wo, Ho, kBlockM, lbw, lbh, lbn);
  const int m_gen = dm::cdiv(Wo, 1 << lbw) * dm::cdiv(Ho, 1 << lbh) * dm::cdiv(N, 1 << lbn);     // tiles of the generic kernels
  const int block_n = halo ? pick_block_n_tiles(Cout, dm::cdiv((long)N * Ho * Wo, kBlockM), g_debug[5] != 2)
                           : pick_block_n_tiles(Cout, m_gen, pair_ok(m_gen));
  fill_common(P, N, Ho, Wo, Cout, block_n);
  P.chunks0 = dm::cdiv(C0, 64);
  P.chunks1 = x1 ? dm::cdiv(C1, 64) : 0;
  P.dual = x1 ? 1 : 0;
  P.num_taps = kh * kw;
  if (halo) {
    P.log_bw = 3; P.log_bh = 4; P.log_bn = 0;
    P.tiles_w = Wo / 8; P.tiles_h = Ho / 16; P.tiles_b = N;
    P.m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
  }
  const int bw = halo ? kHaloW : 1 << P.log_bw, bh = halo ? kHaloH : 1 << P.log_bh, bn = 1 << P.log_bn;
  int rc;
  if (stride == 1) {
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s) { TapInfo t = {(int8_t)(r - pad), (int8_t)(s - pad), 0, 0}; P.taps[r * kw + s] = t; }
    rc = make_act_map(&P.tmA[0], x0, C0, Win, Hin, N, ld0, (long long)Win * ld0, (long long)Hin * Win * ld0, bw, bh, bn);
    if (rc) return rc;
    if (x1) {
      rc = make_act_map(&P.tmA[1], x1, C1, Win, Hin, N, ld1, (long long)Win * ld1, (long long)Hin * Win * ld1, bw, bh, bn);
      if (rc) return rc;
    }
  } else {
    // input row = 2*y - pad + r = 2*(y + floor((r-pad)/2)) + ((r-pad) mod 2): parity view + integer offset
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s) {
        const int a = r - pad, b = s - pad;
        const int pa = ((a % 2) + 2) % 2, pb = ((b % 2) + 2) % 2;
        TapInfo t = {(int8_t)((a - pa) / 2), (int8_t)((b - pb) / 2), (int8_t)(pa * 2 + pb), 0};
        P.taps[r * kw + s] = t;
      }
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb) {
        const bf16* base = reinterpret_cast<const bf16*>(x0) + ((long long)pa * Win + pb) * ld0;
        rc = make_act_map(&P.tmA[pa * 2 + pb], base, C0, Win / 2, Hin / 2, N, 2LL * ld0, 2LL * Win * ld0,
                          (long long)Hin * Win * ld0, bw, bh, bn);
        if (rc) return rc;
      }
  }
  const long long ktot = (long long)P.num_taps * (P.chunks0 + P.chunks1) * 64;
  rc = make_w_map(&P.tmB, wpk, Cout, ktot, block_n);
  if (rc) return rc;
  P.out = y; P.out_f32 = y_f32;
  P.sN = (long long)Ho * Wo * ldy; P.sH = (long long)Wo * ldy; P.sW = ldy;
  P.Cout = Cout; P.ldc_pad = ldy;
  P.bias = bias; P.scale = scale; P.act = act; P.stats = stats; P.stats_ld = stats_ld;
  P.stat_rows = stats ? dm_conv2d_fwd_stat_rows(N, Ho, Wo, Cout) : 0;
  if (stats && (scale || act)) { dm_set_error("dm_conv2d_fwd: statistics are taken of the plain conv output (no scale/act)"); return DM_ERR_ARG; }
  // CTA pairs when the tile width splits into two halves of whole 8-row swizzle groups (always: block_n % 16 == 0)
  if (halo && g_debug[5] != 2 && (block_n % 32) == 0) return launch_conv_halo2(P, wpk, Cout, ktot, (cudaStream_t)stream);
  return halo ? launch_conv_halo(P, (cudaStream_t)stream) : launch_conv(P, (cudaStream_t)stream, wpk, Cout, ktot);
}

// widest conv whose epilogue can take the BatchNorm statistics (the slab must fit beside the operand rings)
extern "C" int dm_conv2d_fwd_stats_max_cout(void) { return kMaxStatBytes / 32 - 256; }

// rows of the statistics buffer dm_conv2d_fwd writes: one per CTA of the persistent grid
extern "C" int dm_conv2d_fwd_stat_rows(int N, int Ho, int Wo, int Cout) {
  // one row per CTA; the two tilings (generic 128-pixel patches / 8x16 halo tiles) can differ in tile
  // count, so this is the larger of the two and the kernels zero-fill the rows beyond their grid
  int a, b, c; pick_patch(Wo, Ho, kBlockM, a, b, c);
  const int mt = dm::cdiv((long)N * Ho * Wo, kBlockM);
  const int bn_a = pick_block_n_tiles(Cout, mt, false), bn_b = pick_block_n_tiles(Cout, mt, true);
  const int nt = dm::cdiv(Cout, bn_a < bn_b ? bn_a : bn_b);          // whichever kernel runs: the larger tile count
  const int m_generic = dm::cdiv(Wo, 1 << a) * dm::cdiv(Ho, 1 << b) * dm::cdiv(N, 1 << c);
  const int m_halo = dm::cdiv(Wo, 8) * dm::cdiv(Ho, 16) * N;
  const int m = m_generic > m_halo ? m_generic : m_halo;
  return conv_grid((m + 1) * nt);       // +1: the CTA-pair kernels round an odd M-tile count up to whole pairs
}

// ---------------------------------------------------------------------------------------------------
// dm_conv2d_s2_dgrad: data gradient of a k=4, stride 2, pad 1 convolution.
// dy [N,Ho,Wo,Cout] -> dx [N,2Ho,2Wo,Cin].  Four launches, one per output parity (ph,pw); each is a
// 2x2-tap conv over dy.  wpk: [4 phases][Cin][4 taps][Cout_k] (see dm_pack_weights mode 2).
extern "C" int dm_conv2d_s2_dgrad(const void* dy, int Cout, int lddy, const void* wpk, void* dx, int Cin, int lddx,
                                  int N, int Ho, int Wo, void* stream) {
  if (ensure_encode() != DM_OK) return DM_ERR_TMA;
  if ((lddy & 7) || (lddx & 7)) { dm_set_error("dm_conv2d_s2_dgrad: channel pitch must be a multiple of 8"); return DM_ERR_ARG; }
  const int H = 2 * Ho, W = 2 * Wo;
  const int chunks = dm::cdiv(Cout, 64);
  const long long ktot = 4LL * chunks * 64;
  ConvParams P;
  memset(&P, 0, sizeof P);
  int lbw, lbh, lbn;
  pick_patch(Wo, Ho, kBlockM, lbw, lbh, lbn);
  const int m_gen = dm::cdiv(Wo, 1 << lbw) * dm::cdiv(Ho, 1 << lbh) * dm::cdiv(N, 1 << lbn);
  const bool pairs = pair_ok(m_gen);
  // the four parities are four phases of ONE launch (4x the tiles for the persistent grid to balance)
  const int block_n = pick_block_n_tiles(Cin, pairs ? 4 * ((m_gen + 1) / 2) * 2 : 4 * m_gen, pairs);
  fill_common(P, N, Ho, Wo, Cin, block_n);
  P.chunks0 = chunks; P.num_taps = 4;
  P.phases = 4; P.phase_rows = Cin;
  P.out_ph = (long long)W * lddx; P.out_pw = lddx;
  // row i = 2Y+ph gets dy rows y with r = i + 1 - 2y in [0,3]:  ph=0: (y=Y, r=1), (y=Y-1, r=3);  ph=1: (y=Y+1, r=0), (y=Y, r=2)
  const int dyo[2][2] = {{0, -1}, {1, 0}};
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw)
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          TapInfo t = {(int8_t)dyo[ph][a], (int8_t)dyo[pw][b], 0, 0};
          P.taps[(ph * 2 + pw) * 4 + a * 2 + b] = t;
        }
  int rc = make_act_map(&P.tmA[0], dy, Cout, Wo, Ho, N, lddy, (long long)Wo * lddy, (long long)Ho * Wo * lddy,
                        1 << P.log_bw, 1 << P.log_bh, 1 << P.log_bn);
  if (rc) return rc;
  rc = make_w_map(&P.tmB, wpk, 4LL * Cin, ktot, block_n);       // [4 phases][Cin] rows
  if (rc) return rc;
  P.out = dx;
  P.out_f32 = 0;
  P.sN = (long long)H * W * lddx; P.sH = 2LL * W * lddx; P.sW = 2LL * lddx;
  P.Cout = Cin; P.ldc_pad = lddx;
  return launch_conv(P, (cudaStream_t)stream, wpk, 4LL * Cin, ktot);
}

// ---------------------------------------------------------------------------------------------------
// dm_convt_fwd: ConvTranspose2d with kernel == stride == k (non-overlapping): a GEMM
// [N*Hin*Win, Cin] x [Cin, k*k*Cout] with a pixel-shuffle scatter epilogue.  wpk: [k*k*Cout][Cin_k].
extern "C" int dm_convt_fwd(const void* x, int Cin, int ldx, const void* wpk, const float* bias, void* y, int ldy, int N,
                            int Hin, int Win, int Cout, int k, void* stream) {
  if (ensure_encode() != DM_OK) return DM_ERR_TMA;
  if ((ldx & 7) || (ldy & 7) || (Cout & 15)) { dm_set_error("dm_convt_fwd: Cout must be a multiple of 16, pitches of 8"); return DM_ERR_ARG; }
  int block_n = 256;
  while (Cout % block_n) block_n -= 16;
  ConvParams P;
  memset(&P, 0, sizeof P);
  fill_common(P, N, Hin, Win, k * k * Cout, block_n);
  P.chunks0 = dm::cdiv(Cin, 64); P.num_taps = 1;
  TapInfo t = {0, 0, 0, 0}; P.taps[0] = t;
  int rc = make_act_map(&P.tmA[0], x, Cin, Win, Hin, N, ldx, (long long)Win * ldx, (long long)Hin * Win * ldx,
                        1 << P.log_bw, 1 << P.log_bh, 1 << P.log_bn);
  if (rc) return rc;
  rc = make_w_map(&P.tmB, wpk, (long long)k * k * Cout, (long long)P.chunks0 * 64, block_n);
  if (rc) return rc;
  const int Ho = Hin * k, Wo = Win * k;
  P.out = y; P.out_f32 = 0;
  P.sN = (long long)Ho * Wo * ldy; P.sH = (long long)k * Wo * ldy; P.sW = (long long)k * ldy;
  P.sKy = (long long)Wo * ldy; P.sKx = ldy; P.convt_k = k;
  P.Cout = Cout; P.ldc_pad = ldy;
  P.bias = bias;
  return launch_conv(P, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// Split-K choice for the weight-gradient GEMMs: `base` output tiles, `patches` 64-pixel K-steps.
// Static round-robin tile schedule => time ~ waves * (K-steps per split + per-tile overhead); pick the
// split count that minimises it (i.e. fills the last wave) instead of a fixed "two waves" rule.
static int pick_splits(int base, int patches, int tile_overhead_steps, int slots = 0) {
  if (slots <= 0) slots = g_sm_limit;
  int max_splits = patches / 8; if (max_splits < 1) max_splits = 1;
  if (max_splits > 1024) max_splits = 1024;
  int best = 1; double best_cost = 1e30;
  for (int s = 1; s <= max_splits; ++s) {
    const int pps = dm::cdiv(patches, s);
    const int se = dm::cdiv(patches, pps);
    if (se != s) continue;
    const long long tiles = (long long)base * se;
    const long long waves = (tiles + slots - 1) / slots;
    const double cost = (double)waves * (pps + tile_overhead_steps);
    if (cost < best_cost * 0.999) { best_cost = cost; best = s; }
  }
  return best;
}

bool g_attr_c = false;

// dm_conv2d_wgrad: dWp[Cout][taps][Cin_k] (fp32, accumulated with red.add) += dY^T * im2col(X).
// Same geometry arguments as dm_conv2d_fwd; dy is [N,Ho,Wo,Cout] bf16 with pitch lddy.
// Two kernels: wgrad_gemm_kernel (M = Cout tiles of 128, for Cout a multiple of 128) and
// wgrad2_gemm_kernel (M = im2col boxes, two accumulators per CTA; everything else).
extern "C" int dm_conv2d_wgrad(const void* x0, int C0, int ld0, const void* x1, int C1, int ld1, const void* dy, int lddy,
                               float* dwp, int N, int Hin, int Win, int Cout, int kh, int kw, int stride, int pad,
                               void* stream) {
  if (ensure_encode() != DM_OK) return DM_ERR_TMA;
  if (kh * kw > 16 || (stride != 1 && stride != 2)) { dm_set_error("dm_conv2d_wgrad: unsupported kernel/stride"); return DM_ERR_ARG; }
  if (stride == 2 && (x1 != nullptr || (Hin & 1) || (Win & 1))) { dm_set_error("dm_conv2d_wgrad: stride 2 needs one source, even H/W"); return DM_ERR_ARG; }
  if ((ld0 & 7) || (x1 && (ld1 & 7)) || (lddy & 7)) { dm_set_error("dm_conv2d_wgrad: channel pitch must be a multiple of 8"); return DM_ERR_ARG; }
  const int Ho = (Hin + 2 * pad - kh) / stride + 1, Wo = (Win + 2 * pad - kw) / stride + 1;
  // ---- geometry shared by both kernels: 64-pixel patches, filter taps, activation maps
  int log_bw, log_bh, log_bn;
  pick_patch(Wo, Ho, 64, log_bw, log_bh, log_bn);
  const int bw = 1 << log_bw, bh = 1 << log_bh, bn = 1 << log_bn;
  const int tiles_w = dm::cdiv(Wo, bw), tiles_h = dm::cdiv(Ho, bh), tiles_b = dm::cdiv(N, bn);
  const int patches = tiles_w * tiles_h * tiles_b;
  const int chunks0 = dm::cdiv(C0, 64), chunks1 = x1 ? dm::cdiv(C1, 64) : 0;
  const int chunks = chunks0 + chunks1;
  const int num_taps = kh * kw;
  TapInfo taps[16];
  CUtensorMap tmX[4], tmDY;
  int rc;
  if (stride == 1) {
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s) { TapInfo t = {(int8_t)(r - pad), (int8_t)(s - pad), 0, 0}; taps[r * kw + s] = t; }
    rc = make_act_map(&tmX[0], x0, C0, Win, Hin, N, ld0, (long long)Win * ld0, (long long)Hin * Win * ld0, bw, bh, bn);
    if (rc) return rc;
    if (x1) {
      rc = make_act_map(&tmX[1], x1, C1, Win, Hin, N, ld1, (long long)Win * ld1, (long long)Hin * Win * ld1, bw, bh, bn);
      if (rc) return rc;
    }
  } else {
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s) {
        const int a = r - pad, b = s - pad;
        const int pa = ((a % 2) + 2) % 2, pb = ((b % 2) + 2) % 2;
        TapInfo t = {(int8_t)((a - pa) / 2), (int8_t)((b - pb) / 2), (int8_t)(pa * 2 + pb), 0};
        taps[r * kw + s] = t;
      }
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb) {
        const bf16* base = reinterpret_cast<const bf16*>(x0) + ((long long)pa * Win + pb) * ld0;
        rc = make_act_map(&tmX[pa * 2 + pb], base, C0, Win / 2, Hin / 2, N, 2LL * ld0, 2LL * Win * ld0,
                          (long long)Hin * Win * ld0, bw, bh, bn);
        if (rc) return rc;
      }
  }
  rc = make_act_map(&tmDY, dy, Cout, Wo, Ho, N, lddy, (long long)Wo * lddy, (long long)Ho * Wo * lddy, bw, bh, bn);
  if (rc) return rc;

  bool use_v2 = (Cout % 128) != 0 || Cout <= 384 || (stride == 2 && Cout <= 768);
  if (num_taps == 1) use_v2 = false;          // 1x1: one box per 64 input channels, kernel B's Cout tiles waste nothing
  if (g_debug[4] == 1) use_v2 = false;
  if (g_debug[4] == 2) use_v2 = true;

  // Kernel F (CTA pairs) is slower than kernel C where C applies: with both operands MN-major an M=256
  // instruction takes ~250 cycles against ~170 for two M=128 ones' worth of work per SM pair (the K-major conv
  // kernel E gets 137), so pairs only win on the short-K small-spatial layers with wide Cout, where kernel B's
  // one accumulator per CTA leaves the MMA pipe idle during its epilogue.
  bool use_pair = !use_v2 && num_taps > 1 && (Cout % 128) == 0 && patches <= 64 && (Cout >= 1024 || chunks >= 32);
  if (g_debug[4] == 3) use_pair = true;
  if (g_debug[4] == 1 || g_debug[4] == 2) use_pair = false;
  if (use_pair) {
    // CTA pairs (kernel F): N tile = whole Cout up to 256, each CTA stages ceil(N/2 / 64) dY boxes
    Wgrad2Params P;
    memset(&P, 0, sizeof P);
    memcpy(P.tmX, tmX, sizeof tmX); P.tmDY = tmDY; memcpy(P.taps, taps, sizeof taps);
    P.num_taps = num_taps; P.chunks0 = chunks0; P.chunks1 = chunks1; P.dual = x1 ? 1 : 0;
    P.log_bw = log_bw; P.log_bh = log_bh; P.log_bn = log_bn;
    P.tiles_w = tiles_w; P.tiles_h = tiles_h; P.tiles_b = tiles_b; P.patches = patches;
    P.total_boxes = num_taps * chunks;
    P.m_tiles = dm::cdiv(P.total_boxes, 4);
    P.n_tiles = dm::cdiv(Cout, 256);
    P.block_n = (dm::cdiv(Cout, P.n_tiles) + 31) / 32 * 32;      // both halves multiples of 16 columns
    P.nb = dm::cdiv(P.block_n / 2, 64);                         // boxes per CTA
    P.Cout = Cout; P.cin_k = chunks * 64; P.C0 = C0; P.C1 = C1;
    int splits = pick_splits(P.m_tiles * P.n_tiles, patches, 3, g_sm_limit / 2);
    if (g_debug[2] > 0) splits = (int)g_debug[2];
    P.patches_per_split = dm::cdiv(patches, splits);
    P.splits = dm::cdiv(patches, P.patches_per_split);
    P.dwp = dwp;
    P.idesc = (make_idesc(P.block_n, true, true) & ~(0x1Fu << 24)) | ((256u >> 4) << 24);
    P.desc_hi = kDescHiMN; P.lbo = 8192 >> 4;
    const int stage_bytes = (2 + P.nb) * 8192;
    int stages = (kSmemBudget - 1024 - 512) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (g_debug[0] > 0 && g_debug[0] < stages) stages = (int)g_debug[0];
    if (g_debug[10] > 0 && g_debug[10] < stages) stages = (int)g_debug[10];   // dev: wgrad-only ring depth (leaves shared memory for co-resident kernels)
    P.stages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 512;
    static bool attr_f = false;
    if (!attr_f) {
      cudaError_t e = cudaFuncSetAttribute(wgrad3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
      attr_f = true;
    }
    const int tiles = P.m_tiles * P.n_tiles * P.splits;
    const int clusters = tiles < g_sm_limit / 2 ? tiles : g_sm_limit / 2;
    dm_note_kernel("wgrad3_pair", P.splits);
    wgrad3_pair_kernel<<<2 * clusters, kThreads, smem, (cudaStream_t)stream>>>(P);
    DM_CHECK_LAUNCH();
    return DM_OK;
  }
  if (use_v2) {
    Wgrad2Params P;
    memset(&P, 0, sizeof P);
    memcpy(P.tmX, tmX, sizeof tmX); P.tmDY = tmDY; memcpy(P.taps, taps, sizeof taps);
    P.num_taps = num_taps; P.chunks0 = chunks0; P.chunks1 = chunks1; P.dual = x1 ? 1 : 0;
    P.log_bw = log_bw; P.log_bh = log_bh; P.log_bn = log_bn;
    P.tiles_w = tiles_w; P.tiles_h = tiles_h; P.tiles_b = tiles_b; P.patches = patches;
    P.total_boxes = num_taps * chunks;
    P.m_tiles = dm::cdiv(P.total_boxes, 4);
    P.n_tiles = dm::cdiv(Cout, 256);
    P.nb = dm::cdiv(dm::cdiv(Cout, P.n_tiles), 64);
    P.block_n = P.nb * 64;
    P.Cout = Cout; P.cin_k = chunks * 64; P.C0 = C0; P.C1 = C1;
    int splits = pick_splits(P.m_tiles * P.n_tiles, patches, 8);
    if (g_debug[2] > 0) splits = (int)g_debug[2];
    P.patches_per_split = dm::cdiv(patches, splits);
    P.splits = dm::cdiv(patches, P.patches_per_split);
    P.dwp = dwp;
    P.idesc = make_idesc(P.block_n, true, true);
    P.desc_hi = kDescHiMN; P.lbo = 8192 >> 4;
    const int stage_bytes = (4 + P.nb) * 8192;
    int stages = (kSmemBudget - 1024 - 512) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (g_debug[0] > 0 && g_debug[0] < stages) stages = (int)g_debug[0];
    if (g_debug[10] > 0 && g_debug[10] < stages) stages = (int)g_debug[10];   // dev: wgrad-only ring depth (leaves shared memory for co-resident kernels)
    P.stages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 512;
    if (!g_attr_c) {
      cudaError_t e = cudaFuncSetAttribute(wgrad2_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
      if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
      g_attr_c = true;
    }
    const int tiles = P.m_tiles * P.n_tiles * P.splits;
    int grid = tiles < g_sm_limit ? tiles : g_sm_limit;
    if (g_debug[1] > 0 && g_debug[1] < grid) grid = (int)g_debug[1];
    dm_note_kernel("wgrad2_gemm", P.splits);
    wgrad2_gemm_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(P);
    DM_CHECK_LAUNCH();
    return DM_OK;
  }

  WgradParams P;
  memset(&P, 0, sizeof P);
  memcpy(P.tmX, tmX, sizeof tmX); P.tmDY = tmDY; memcpy(P.taps, taps, sizeof taps);
  P.log_bw = log_bw; P.log_bh = log_bh; P.log_bn = log_bn;
  P.tiles_w = tiles_w; P.tiles_h = tiles_h; P.tiles_b = tiles_b; P.patches = patches;
  P.chunks0 = chunks0; P.chunks1 = chunks1; P.dual = x1 ? 1 : 0;
  P.num_taps = num_taps;
  P.cin_k = chunks * 64;
  P.block_n = chunks >= 4 ? 256 : chunks * 64;
  P.ci_tiles = dm::cdiv(P.cin_k, P.block_n);
  P.co_tiles = dm::cdiv(Cout, 128);
  P.Cout = Cout;
  const int base_tiles = P.co_tiles * P.num_taps * P.ci_tiles;
  int splits = pick_splits(base_tiles, patches, 3);
  if (g_debug[2] > 0) splits = (int)g_debug[2];
  P.patches_per_split = dm::cdiv(P.patches, splits);
  P.splits = dm::cdiv(P.patches, P.patches_per_split);
  P.b_stage_bytes = (P.block_n / 64) * 8192;
  P.dwp = dwp;
  P.idesc = make_idesc(P.block_n, true, true);
  P.desc_hi_a = kDescHiMN; P.desc_hi_b = kDescHiMN;
  P.lbo_a = 8192 >> 4; P.lbo_b = 8192 >> 4;
  int stage_bytes = kAStage + P.b_stage_bytes;
  int stages = (kSmemBudget - 1024 - kAuxBytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (g_debug[0] > 0 && g_debug[0] < stages) stages = (int)g_debug[0];
    if (g_debug[10] > 0 && g_debug[10] < stages) stages = (int)g_debug[10];   // dev: wgrad-only ring depth (leaves shared memory for co-resident kernels)
  P.stages = stages;
  size_t smem = 1024 + (size_t)stages * stage_bytes + kAuxBytes;
  if (!g_attr_b) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) { dm_set_error(cudaGetErrorString(e)); return DM_ERR_CUDA; }
    g_attr_b = true;
  }
  int tiles = base_tiles * P.splits;
  int grid = tiles < g_sm_limit ? tiles : g_sm_limit;
  if (g_debug[1] > 0 && g_debug[1] < grid) grid = (int)g_debug[1];
  dm_note_kernel("wgrad_gemm", P.splits);
  wgrad_gemm_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(P);
  DM_CHECK_LAUNCH();
  return DM_OK;
}
