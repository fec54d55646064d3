/* dm_b200.h -- C-ABI of libdm_b200.so: the B200 (sm_100a) kernels behind the DiffusionModel hot path.
 *
 * The reference (Shen-Yuuu/DiffusionModel) has no FFI layer: its hot path is torch.nn modules calling
 * aten -> cuDNN/cuBLAS.  The boundary this library replaces is therefore the set of aten/cuDNN calls
 * made by ContextUnet / DDPM (file:line under /root/reference cited per entry point); the Python host
 * mirror (diffusionmodel_b200/unet.py, ddpm.py) keeps the reference's module names, signatures and
 * state_dict layout and binds these symbols through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; `stream` is a cudaStream_t passed as void*;
 *   - activations are bf16, NHWC, with a channel pitch `ld*` (elements, multiple of 8; pad lanes are
 *     written as zero); "P" is a pixel count (N*H*W); small per-(sample,channel) tensors are fp32;
 *   - parameter-gradient outputs are ACCUMULATED (+=) into fp32 buffers in the reference's layouts;
 *   - every function returns 0 on success, <0 on error (dm_last_error() has the message); nothing
 *     falls back to the CPU and nothing synchronises the stream.
 */
#ifndef DM_B200_H_
#define DM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* dm_last_error(void);
int dm_version(void);
void dm_debug_set(int key, long long value);
long long dm_launch_count(void);   /* kernels launched by this library so far (bench.py's gpu_launches) */
/* Which kernel variant the conv / weight-gradient dispatchers chose (tests assert the selection at the benchmarked
 * shapes): launches so far under a kernel name ("conv_gemm", "conv3x3_halo", "conv3x3_halo2", "wgrad_gemm",
 * "wgrad2_gemm", "wgrad3_pair", "skinny_gemm"), and the most recent launch with its parameter (conv: tile width in
 * output channels; wgrad: split-K count). */
/* SMs the persistent convolution / weight-gradient grids may occupy (default and maximum 148; rounded down to even).
 * Lower it while another stream runs a long kernel that also needs SM residency (the data-parallel gradient all-reduce
 * overlapped with the encoder's backward pass): a GEMM CTA owns a whole SM's shared memory and cannot share one.  The value
 * is read at launch time, so a CUDA graph captured under a limit keeps it. */
int dm_set_sm_limit(int sms);
long long dm_kernel_count(const char* name);
const char* dm_last_kernel(int* param);
/* Device scratch (>= 16 MiB recommended) for the partial sums of the reduction kernels (pooling, colsum, FiLM
 * gradients): used in stream order by every call, so one buffer per stream; must outlive captured graphs. */
int dm_set_workspace(void* ptr, long long bytes);

/* ---- tcgen05 implicit-GEMM convolutions (conv_gemm.cu) ------------------------------------------
 * nn.Conv2d forward: new_scripy.py:166,169,184,188,217,222,225,229,243,311,314; MNIST_script.py:42,46,149,152.
 * x0 (++ x1 concatenated on channels: replaces torch.cat new_scripy.py:355 / MNIST_script.py:186),
 * wpk = dm_pack_weight() output [Cout][kh*kw][Cin_k]; y bf16 (y_f32=0) or fp32 (y_f32=1, ldy%4==0);
 * stats (nullable): [dm_conv2d_fwd_stat_rows()][2][stats_ld] per-CTA sum / sum-of-squares of y for the
 * train-mode BatchNorm that follows: of the values as stored (bf16-rounded), accumulated across the CTA's tiles in
 * shared memory in a fixed order (bit-reproducible).
 * y = act(conv * scale + bias): scale (nullable) and act (0 none, 1 GELU, 2 ReLU) fold an EVAL-mode
 * BatchNorm2d + activation into the epilogue (sampling loop: running statistics are constants).  The data gradient of a stride-1 conv is the same call with the
 * flipped/transposed weight pack and pad' = k-1-pad. */
int dm_conv2d_fwd(const void* x0, int C0, int ld0, const void* x1, int C1, int ld1, const void* wpk,
                  const float* bias, const float* scale, int act, void* y, int ldy, int y_f32, float* stats,
                  int stats_ld, int N, int Hin, int Win, int Cout, int kh, int kw, int stride, int pad, void* stream);
int dm_conv2d_fwd_stat_rows(int N, int Ho, int Wo, int Cout);
/* Widest Cout for which dm_conv2d_fwd accepts `stats` (the per-CTA slab must fit beside the operand rings). */
int dm_conv2d_fwd_stats_max_cout(void);
/* data gradient of Conv2d(k=4, s=2, p=1) (new_scripy.py:229); wpk = 4 phase packs, see weights.py */
int dm_conv2d_s2_dgrad(const void* dy, int Cout, int lddy, const void* wpk, void* dx, int Cin, int lddx,
                       int N, int Ho, int Wo, void* stream);
/* nn.ConvTranspose2d with kernel == stride (new_scripy.py:298; MNIST_script.py:88,141) */
int dm_convt_fwd(const void* x, int Cin, int ldx, const void* wpk, const float* bias, void* y, int ldy, int N,
                 int Hin, int Win, int Cout, int k, void* stream);
/* weight gradient: dwp[Cout][kh*kw][Cin_k] fp32 += dy^T * im2col(x) (cuDNN wgrad of the sites above) */
int dm_conv2d_wgrad(const void* x0, int C0, int ld0, const void* x1, int C1, int ld1, const void* dy, int lddy,
                    float* dwp, int N, int Hin, int Win, int Cout, int kh, int kw, int stride, int pad,
                    void* stream);

/* ---- weight packing (pack.cu) --------------------------------------------------------------------
 * out[row][tap][col] (tap_major_rows=0) or out[tap][row][col] (=1) in bf16, col padded with zeros to
 * cols_k, row padded to row_len; source element = w[row*s_row + col*s_col + tap_off[tap]].
 * c_split>0: columns >= c_split start at round_up(c_split,64) (dual-source convs). */
int dm_pack_weight(const float* w, void* out, int rows, int cols, int ntaps, const long long* tap_off_host,
                   long long s_row, long long s_col, int c_split, int cols_k, long long row_len,
                   int tap_major_rows, void* stream);
/* inverse scatter: grad[...] += dwp[...] with the same addressing (fp32 -> fp32); consume != 0 also
 * re-zeroes the packed accumulator (persistent per-parameter accumulators, flushed once per optimizer step) */
/* data-gradient packs of a weight whose bf16 forward pack is [Cout][ntaps][Cin]: per tap a tiled transpose to
 * dst[ci*dst_pitch + dst_off[tap] + co], co < ceil64(Cout) (zeros beyond Cout); replaces the strided gather of
 * dm_pack_weight for weights stored GEMM-natively by the optimizer */
int dm_pack_transpose(const void* src, void* dst, int cout, int cin, int ntaps, const long long* dst_off,
                      long long dst_pitch, void* stream);
int dm_unpack_wgrad(float* dwp, float* grad, int rows, int cols, int ntaps, const long long* tap_off_host,
                    long long s_row, long long s_col, int c_split, int cols_k, long long row_len,
                    int tap_major_rows, int consume, void* stream);

/* ---- layout conversion --------------------------------------------------------------------------- */
int dm_nchw_to_nhwc(const float* x, void* y, int ldy, int y_f32, int N, int C, int H, int W, void* stream);
int dm_cast_nhwc(const float* x, int ldx, void* y, int ldy, long long P, int C, void* stream);  /* fp32 -> bf16 NHWC */
int dm_nhwc_to_nchw(const void* x, int x_f32, int ldx, float* y, int N, int C, int H, int W, void* stream);
/* 3x3/pad-1 im2col of a <=3-channel image into [N,H,W,32] (column ci*9 + r*3 + s): the first U-Net conv
 * (new_scripy.py:184 with in_ch=3, MNIST_script.py:42 with 1) then runs as a 1x1 convolution */
int dm_im2col3x3(const void* x, int ldx, void* out, int N, int H, int W, int C, void* stream);
int dm_space_to_depth(const void* x, int ldx, void* y, int ldy, int N, int H, int W, int C, int k, int chan_major,
                      void* stream);   /* out channel = tap*C + c, or c*k*k + tap when chan_major */

/* ---- BatchNorm2d + GELU (new_scripy.py:185-186,189-190,218-219,226-227) ---------------------------
 * finalize: partial sums -> batch mean / invstd (+ running-stat update, momentum 0.1, unbiased var);
 * m_tiles == 0: eval mode, mean/invstd from the running buffers. */
/* batch statistics of a stored conv output: part[dm_bn_stats_rows(P, C)][2][ld] partial sums of y, y^2
 * (no atomics, fixed order: bit-reproducible); alternative to the conv epilogue's fused statistics */
int dm_bn_stats_rows(long long P, int C);
int dm_bn_stats(const void* y, int ldy, float* part, int ld, long long P, int C, void* stream);
int dm_bn_finalize(const float* partials, int m_tiles, int ld, int C, double count, float* mean, float* invstd,
                   float* running_mean, float* running_var, float momentum, float eps, const float* conv_bias,
                   void* stream);   /* conv_bias (nullable): bias the producing conv left out of y (train mode: the
                                       norm cancels it); it is added to the tracked running mean only */
int dm_bn_act_fwd(const void* y, int ldy, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, void* z, int ldz, long long P, int C, int act, void* stream);
/* dy = BN/act backward (three launches: block partial sums -> finalize -> apply); dgamma/dbeta += ;
 * dbias (nullable) += the bias gradient of the convolution that feeds this norm (sum_p dy, evaluated
 * from the same sums instead of a separate pass over dy); scratch >= dm_bn_act_bwd_scratch(P, C)
 * floats; training=0 drops the batch-statistic terms. */
long long dm_bn_act_bwd_scratch(long long P, int C);
int dm_bn_act_bwd(const void* dz, int lddz, const void* y, int ldy, const float* mean, const float* invstd,
                  const float* gamma, const float* beta, void* dy, int lddy, float* dgamma, float* dbeta,
                  float* dbias, float* scratch, long long P, int C, int act, int training, void* stream);

/* ---- GroupNorm(8,C) + ReLU/GELU (new_scripy.py:167-168,299-300,312-313) ---------------------------
 * Three launches each way (per-sample partial sums -> fold to per-(sample,channel) coefficient rows ->
 * streaming apply).  scratch >= dm_gn_scratch(N, HW, C) floats; the backward call must receive the SAME
 * scratch buffer the forward call filled (it holds the coefficient rows). */
long long dm_gn_scratch(int N, int HW, int C);
int dm_gn_act_fwd(const void* x, int ldx, const float* gamma, const float* beta, void* z, int ldz, float* mean,
                  float* rstd, float* scratch, int N, int HW, int C, int G, float eps, int act, void* stream);
int dm_gn_act_bwd(const void* dz, int lddz, const void* x, int ldx, const float* mean, const float* rstd,
                  const float* gamma, const float* beta, void* dx, int lddx, float* dgamma, float* dbeta,
                  float* scratch, int N, int HW, int C, int G, int act, void* stream);

/* ---- per-(sample,channel) reductions and the SEBlock gate (new_scripy.py:154-158,196-205) ---------- */
int dm_pool_nhw(const void* x, int ldx, float* out, int N, int HW, int C, float scale, void* stream);
int dm_pool_prod_nhw(const void* a, int lda, const void* b, int ldb, float* out, int N, int HW, int C, float scale,
                     void* stream);
/* out = (res + x2*gate[n,c]) * scale   (gate NULL -> 1) */
int dm_se_apply_fwd(const void* x2, int ld2, const float* gate, const void* res, int ldr, void* out, int ldo,
                    int N, int HW, int C, float scale, void* stream);
/* dx2 = dout*scale*gate + dpool[n,c]/HW ; dres = dout*scale  (gate/dpool nullable) */
int dm_se_apply_bwd(const void* dout, int lddo, const float* gate, const float* dpool, void* dx2, int lddx2,
                    void* dres, int lddr, int N, int HW, int C, float scale, void* stream);

/* ---- skinny GEMM: out[m][n] = sum_k A[m][k] * W[n][k], bf16 operands, M <= 16 rows, K a multiple of 32 -----------
 * The data gradient of up0 = ConvTranspose2d(8F, 8F, 8, 8) on the 2x2 bottleneck (new_scripy.py:297-301): 16 pixels x
 * K = 98304 x N = 1536.  K is split over the grid (2048 per block); scratch = dm_skinny_gemm_scratch(N, K) floats of
 * per-slice partial sums, folded in a fixed order and rounded to bf16 into out (row pitch ldo). */
long long dm_skinny_gemm_scratch(int N, int K);
int dm_skinny_gemm(const void* A, long long lda, const void* W, long long ldw, void* out, int ldo, float* scratch, int M,
                   int N, int K, void* stream);

/* ---- [N, C]-sized fp32 linear layers: SEBlock.fc (new_scripy.py:148-152), EmbedFC.model (new_scripy.py:259-263) ---
 * act: 0 none, 1 GELU (erf), 2 ReLU, 3 sigmoid.  W is the nn.Linear weight [Cout][Cin]; b nullable.
 * fwd: y = act(x W^T + b); pre (nullable) receives the pre-activation (what the GELU/ReLU backward needs).
 * bwd: g = (sum of the nparts rows-blocks dy[p][N][Cout]) * act'(aux), aux = pre (GELU/ReLU) or y (sigmoid);
 *      dW += g^T x, db += column sums of g (both nullable), dx_parts[s][N][Cin] = g[:, slice s] W[slice s, :] for the
 *      dm_linear_bwd_parts(Cout) output-feature slices (nullable).  The next dm_linear_act_bwd in the chain consumes
 *      the partial rows directly (its dy/nparts); dm_sum_parts folds them when a plain tensor is needed. */
int dm_linear_act_fwd(const float* x, const float* W, const float* b, float* pre, float* y, int N, int Cin, int Cout,
                      int act, void* stream);
/* masked one-hot class rows out[N][ncls] (fp32) feeding the context EmbedFCs: one_hot(c) * mask (new_scripy.py:337-340),
 * flip != 0: one_hot(c) * -(1 - mask) (MNIST_script.py:165-171).  c int64 [N]; mask fp32 [N], or int64 when mask_i64 != 0. */
int dm_ctx_onehot(const long long* c, const void* mask, int mask_i64, float* out, int N, int ncls, int flip, void* stream);
int dm_linear_bwd_parts(int Cout);
int dm_linear_act_bwd(const float* dy, int nparts, const float* aux, int act, const float* x, const float* W, float* dW,
                      float* db, float* dx_parts, int N, int Cin, int Cout, void* stream);
int dm_sum_parts(const float* parts, int nparts, float* out, long long n, void* stream);

/* ---- CoordAttn directional pooling and gating (new_scripy.py:97-140) ------------------------------- */
int dm_ca_pool(const void* a, int lda, const void* b, int ldb, float* oh, float* ow, int N, int H, int W, int C,
               float scale_h, float scale_w, void* stream);   /* oh[n,h,c]=scale_h*sum_w a*b ; b NULL -> a */
int dm_ca_gate_fwd(const void* x, int ldx, const float* ah, const float* aw, void* out, int ldo, int N, int H,
                   int W, int C, void* stream);
int dm_ca_gate_bwd(const void* dout, int lddo, const float* ah, const float* aw, const float* dxh,
                   const float* dxw, void* dx, int lddx, int N, int H, int W, int C, void* stream);

/* ---- cat + bilinear x2 (align_corners=True) + FiLM (new_scripy.py:242,251,348-349) ----------------- */
int dm_upcat_fwd(const void* a, int lda, int Ca, const void* b, int ldb, int Cb, void* out, int ldo, int N, int h,
                 int w, void* stream);
/* same with a skip tensor b of Nb samples shared cyclically by the N = k * Nb samples of a (sample n reads b[n % Nb]): the
 * classifier-free-guidance halves of the sampling batch share the encoder's skip tensors (new_scripy.py:463-467 repeats
 * the batch; the encoder sees neither c nor the context mask), so they are never duplicated in memory.  Forward only. */
int dm_upcat_fwd_shared(const void* a, int lda, int Ca, const void* b, int ldb, int Cb, int Nb, void* out, int ldo, int N,
                        int h, int w, void* stream);
int dm_upcat_bwd(const void* dout, int lddo, void* da, int ldda, int Ca, void* db, int lddb, int Cb, int N, int h,
                 int w, void* stream);
int dm_film_fwd(const void* x, int ldx, const float* ce, const float* te, void* out, int ldo, int N, int HW, int C,
                void* stream);
int dm_film_bwd(const void* dout, int lddo, const void* x, int ldx, const float* ce, void* dx, int lddx, float* dce,
                float* dte, int N, int HW, int C, void* stream);

/* ---- pooling (new_scripy.py:290; MNIST_script.py:74,132) ------------------------------------------ */
int dm_avgpool_act_fwd(const void* x, int ldx, void* out, int ldo, int N, int H, int W, int C, int k, int act,
                       void* stream);
int dm_avgpool_act_bwd(const void* dout, int lddo, const void* x, int ldx, void* dx, int lddx, int N, int H, int W,
                       int C, int k, int act, void* stream);
int dm_maxpool2_fwd(const void* x, int ldx, void* out, int ldo, int N, int H, int W, int C, void* stream);
int dm_maxpool2_bwd(const void* dout, int lddo, const void* x, int ldx, void* dx, int lddx, int N, int H, int W,
                    int C, void* stream);

/* ---- LocalEnhancer mask weighting (new_scripy.py:172-174) and generic glue -------------------------- */
int dm_mask_fma(const void* x, int ldx, const void* y, int ldy, const float* mask, float thresh, void* out,
                int ldo, long long P, int C, void* stream);   /* out = (x?x:0) + y*(mask>thresh) */
int dm_axpby(const void* a, int lda, const void* b, int ldb, void* out, int ldo, long long P, int C, float sa,
             float sb, void* stream);                         /* out = sa*a + sb*b (b nullable) */
int dm_colsum(const void* dy, int lddy, float* db, long long P, int C, void* stream);   /* db[c] += sum_p dy */

/* ---- DDPM q-sample, losses, CFG reverse step (new_scripy.py:405-437,467-475; MNIST_script.py:239-252,287-295) */
int dm_q_sample(const float* x, const float* noise, const float* sqrtab, const float* sqrtmab, const long long* ts,
                void* xt, int ldo, int N, int C, int H, int W, void* stream);
/* loss = mean((noise-pred)^2 * w(mask)) + fcw*mean(|pred*h - noise*h|); mask NULL -> plain MSE.
 * pred: fp32 NHWC pitch ldp; noise fp32 NCHW; loss: one float (overwritten); scratch >= 1024 floats. */
int dm_ddpm_loss_fwd(const float* pred, int ldp, const float* noise, const float* mask, float* loss, float* scratch,
                     int N, int C, int H, int W, float hi_t, float mid_t, float hi_w, float mid_w, float lo_w,
                     float fcw, void* stream);
/* dpred: fp32 NHWC, pitch lddp (the gradient of the fp32 prediction) */
int dm_ddpm_loss_bwd(const float* pred, int ldp, const float* noise, const float* mask, const float* gout, void* dpred,
                     int lddp, int N, int C, int H, int W, float hi_t, float mid_t, float hi_w, float mid_w,
                     float lo_w, float fcw, void* stream);
/* eps = (1+w)*eps[:n] - w*eps[n:];  x' = a*(x - eps*b) + s*z;  writes x' (fp32 NCHW) and the doubled
 * bf16 NHWC batch the next U-Net call consumes.  eps: fp32 NHWC [2n,H,W,ldp]; z nullable (i == 1). */
int dm_cfg_reverse_step(const float* eps, int ldp, const float* x, const float* z, float* x_out, void* xt_next,
                        int ldo, float guide_w, float oneover_sqrta, float mab_over_sqrtmab, float sqrt_beta, int n,
                        int C, int H, int W, void* stream);

/* same step with (guide_w, oneover_sqrta, mab_over_sqrtmab, sqrt_beta) read from a 4-float DEVICE buffer, so one
 * captured CUDA graph of the reverse step serves all n_T iterations (z must be a valid buffer: zeros at i == 1) */
int dm_cfg_reverse_step_dev(const float* eps, int ldp, const float* x, const float* z, float* x_out, void* xt_next,
                            int ldo, const float* coef4, int n, int C, int H, int W, void* stream);
/* same with ONE GUIDANCE SCALE PER TRAJECTORY: w[n] (device) replaces coef4[0], so the reference's sequential loop over
 * --guide_scales (new_scripy.py:1036-1041) runs as one trajectory batch */
int dm_cfg_reverse_step_w(const float* eps, int ldp, const float* x, const float* z, float* x_out, void* xt_next,
                          int ldo, const float* coef4, const float* w, int n, int C, int H, int W, void* stream);

/* ---- CoordAttn gate network (new_scripy.py:97-140): everything between the directional pooling and the gating pass,
 * forward in two launches and backward in six (+ one memset).  R = N*L rows per direction (H == W == L), C channels, m = C/16.
 * All tensors fp32 and contiguous: xh/xw/ah/aw [R][C]; conv1_* weight [m][C]; h2w/w2h projections [m][m]; conv_h/conv_w
 * weight [C][m]; u/t/dh [2][R][m]; part [nblk][2][2][m] with nblk = ceil(R / dm_ca_gates_rows_per_block()); stat [2][2][m]
 * (mean, rstd per direction: written by the forward, read by the backward).  training != 0: batch statistics and the
 * running-statistics update of both BatchNorms (momentum, unbiased variance); 0: running statistics. */
typedef struct DmCaGates {
  const float *xh, *xw;
  const float *w1_h, *w1_w, *b1_h, *b1_w;
  const float *bn_g_h, *bn_g_w, *bn_b_h, *bn_b_w;
  float *bn_rm_h, *bn_rm_w, *bn_rv_h, *bn_rv_w;
  const float *wp_h2w, *wp_w2h, *bp_h2w, *bp_w2h;
  const float *wc_h, *wc_w, *bc_h, *bc_w;
  const float *gamma_h, *gamma_w, *alpha, *beta;
  float *u, *part, *stat, *t, *t0, *ah, *aw;          /* u, t, t0 [2][R][m] */
  int R, C, m, nblk, training;
  float eps, momentum;
} DmCaGates;
/* gradients: d_ah/d_aw in, d_xh/d_xw out, g_* accumulated (+=) with fp32 atomics; scal = 4 zero-initialised floats the
 * backward re-arms itself; part as in the forward (separate buffer) */
typedef struct DmCaGatesGrad {
  const float *d_ah, *d_aw;
  float *d_xh, *d_xw, *dh, *dt, *du, *dz, *part, *scal;          /* dh, dt, du [2][R][m], dz [2][R][C]: work buffers */
  float *g_w1_h, *g_w1_w, *g_b1_h, *g_b1_w;
  float *g_bn_g_h, *g_bn_g_w, *g_bn_b_h, *g_bn_b_w;
  float *g_wp_h2w, *g_wp_w2h, *g_bp_h2w, *g_bp_w2h;
  float *g_wc_h, *g_wc_w, *g_bc_h, *g_bc_w;
  float *g_gamma_h, *g_gamma_w, *g_alpha, *g_beta;
} DmCaGatesGrad;
int dm_ca_gates_rows_per_block(void);
int dm_ca_gates_fwd(const DmCaGates* p, void* stream);
int dm_ca_gates_bwd(const DmCaGates* p, const DmCaGatesGrad* q, void* stream);

/* ---- input pipeline tail (next-row 3 of SURVEY 8f; CrackDataset.__getitem__ new_scripy.py:516-551 + transforms :683-688)
 * img_u8 [B][H][W][3] decoded + resized (cached on the device), flip[B] (0/1: RandomHorizontalFlip decisions), box[B][4] =
 * (xmin, ymin, xmax, ymax) already scaled + clamped like :542-545 -> x fp32 [B][3][H][W] = ((u8/255) - mean)/std and
 * mask fp32 [B][H][W] = low | mid (rows >= H/2) | high (rows [ymin,ymax), cols [xmin,xmax); not flipped, as in the reference) */
int dm_prep_batch(const void* img_u8, const int* flip, const int* box, float* x, float* mask, int B, int H, int W,
                  float mean, float stdv, float low, float mid, float high, void* stream);

/* ---- sample quality (next-row 4 of SURVEY 8f): ImageMetrics.calc_ssim / calc_psnr (new_scripy.py:1189-1251) for N pairs of
 * fp32 images [N][elems] (any range: mapped with (x+1)/2 when an image's minimum is negative, like the reference);
 * out [N][2] = (ssim, psnr).  FID (:1146-1187) needs pretrained Inception weights and stays out of scope. */
int dm_image_metrics(const float* a, const float* b, float* out, int N, long long elems, void* stream);

/* ---- optimizer side (new_scripy.py:797-803) --------------------------------------------------------- */
int dm_sumsq(const float* g, long long n, float* out, void* stream);          /* *out = sum g^2, deterministic (fixed-order combine: bit-identical on every data-parallel rank) */
int dm_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
             float eps, float wd, float bc1, float bc2, const float* gnorm_sq, float max_norm, void* stream);
/* dm_adamw that also writes a bf16 copy of the updated parameters (the forward weight pack of convolution weights
 * the optimizer stores in [Cout][tap][Cin] order); n % 4 == 0. */
int dm_adamw_bf16(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1,
                  float beta2, float eps, float wd, float bc1, float bc2, const float* gnorm_sq, float max_norm,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DM_B200_H_ */
