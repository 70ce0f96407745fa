/*
 * tactilesr_b200 -- C ABI of libtactilesr_b200.so (sm_100a only).
 *
 * The reference (wmtlab/tactileSR) has no FFI: its hot path is `nn.Module.forward` + autograd calling ATen / cuDNN /
 * cuBLAS.  The drop-in boundary is therefore the Python module protocol (the modules under tactilesr_b200/model/ keep the reference's
 * classes, constructor arguments and state_dict layout); *underneath* it every operation of the path is one of the
 * entry points below.  Each declaration cites the reference call it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; the library never allocates or frees, never synchronises, and
 *     launches on the given `stream` only;
 *   - activations are NHWC: `x` points at the first of `C` consecutive channels of pixel 0 and `ld` is the channel
 *     count of the underlying buffer (so channel slices of a concat buffer are addressed without a copy);
 *   - `*_bf16` arguments are storage-type codes of an activation operand: 0 = float, 1 = __nv_bfloat16, 2 = __half
 *     (for tsr_bn_backward / tsr_relu_backward, whose operands mix gradients and saved activations: 0 = all float,
 *     1 = all bf16, 2 = gradients bf16 + saved activation fp16 -- the "fp16" precision mode);
 *   - weights and their gradients keep PyTorch's OIHW fp32 layout (`state_dict()` compatible); packed copies are made by
 *     tsr_pack_conv_weight_*;
 *   - return value 0 = success; otherwise an error code (1 bad argument, 2 CUDA error, 3 unsupported, 4 workspace too
 *     small) with a message in tsr_last_error().  There is no CPU fallback and no library (cuDNN/cuBLAS) fallback.
 *   - `*_workspace` functions return the scratch bytes the matching call needs.
 */
#ifndef TACTILESR_B200_H
#define TACTILESR_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tsr_stream_t; /* == cudaStream_t */

/* ---- library ------------------------------------------------------------------------------------------------- */
const char* tsr_last_error(void);
int tsr_version(void);
int tsr_check_device(void);               /* 0 iff the current device is compute capability 10.x */
/* "fp16" precision mode guard: a caller-owned device int that kernels storing fp16 activations set to 1 when a stored value is
   not finite (|x| > 65504 or NaN); NULL switches the check off.  Read it at logging cadence, not per step. */
void tsr_set_f16_overflow_flag(int* flag_dev);
long long tsr_launch_count(void);         /* kernels launched by this library since the last reset */
void tsr_launch_count_reset(void);
void tsr_launch_count_add(long long n);   /* account n launches replayed from a captured CUDA graph */

/* ---- fp32-accurate convolutions (FFMA implicit GEMM) -------------------------------------------------------- */
/* nn.Conv2d weights (model/tactileSR_model.py:41,47,53,168,174,180,186,191,219,220) -> [tap][ci][co] (forward) and
 * [flipped tap][co][ci] (data gradient); either output may be NULL. */
int tsr_pack_conv_weight_f32(const float* w_oihw, float* w_fwd, float* w_dgrad, int Cout, int Cin, int KS,
                             tsr_stream_t stream);
/* out = [relu]( conv_KSxKS(in, w_packed) [+ bias] [+ residual] ), same padding.  Forward of nn.Conv2d (+ the fused
 * `output += x; relu` of MSRB.forward :204-206 and ResBlock.forward :222-225); with w_dgrad it is the data gradient.
 * flags bit0 = ReLU.  Cin % 16 == 0, Cout % 64 == 0, KS in {1,3,5}. */
int tsr_conv2d_f32(const float* in, int in_ld, const float* w_packed, const float* bias, const float* residual,
                   int res_ld, float* out, int out_ld, int B, int H, int W, int Cin, int Cout, int KS, int flags,
                   tsr_stream_t stream);
size_t tsr_conv2d_wgrad_f32_workspace(int B, int H, int W, int Cin, int Cout, int KS);
/* dw_oihw (+)= sum_pix in[pix+shift] (x) dout[pix]: autograd weight gradient of nn.Conv2d (cpu/trainer.py:353);
 * deterministic two-level (split-K, then fixed-order) reduction. */
int tsr_conv2d_wgrad_f32(const float* in, int in_ld, const float* dout, int dout_ld, float* dw_oihw, void* workspace,
                         size_t ws_bytes, int B, int H, int W, int Cin, int Cout, int KS, int accumulate,
                         tsr_stream_t stream);
size_t tsr_colsum_workspace(long long npix, int C);
/* out[c] (+)= sum_pix x[pix][c]: bias gradient of nn.Conv2d. */
int tsr_colsum(const void* x, int ld, int x_bf16, long long npix, int C, float* out, void* workspace, size_t ws_bytes,
               int accumulate, tsr_stream_t stream);

/* ---- head / tail convolutions (memory bound) ------------------------------------------------------------------ */
/* nn.Upsample(scale_factor=sf, bilinear, align_corners=False) + nn.Conv2d(3 -> 64, 3x3, no bias) [+ ReLU]
 * (model/tactileSR_model.py:35-37, 60-62; TactileSRCNN :107 + :122).  x: NCHW fp32 (B, *, 4, 4), 3 channels starting
 * at `x`, sample stride x_bstride floats. */
int tsr_head_fwd(const float* x, long long x_bstride, const float* w_oihw, void* out, int out_ld, int out_bf16, int B,
                 int sf, int relu, tsr_stream_t stream);
size_t tsr_head_wgrad_workspace(int B);
/* d(loss)/d(w) of the head convolution from the gradient of its output (autograd of :37 / :61; cpu/trainer.py:353).
 * dout: NHWC, 64 channels starting at `dout`, row stride dout_ld elements (a multiple of 8: the rows are staged by 16-byte
 * asynchronous copies), storage type dout_bf16 (0 fp32, 1 bf16, 2 fp16).  Deterministic two-level reduction. */
int tsr_head_wgrad(const float* x, long long x_bstride, const void* dout, int dout_ld, int dout_bf16, float* dw_oihw,
                   void* workspace, size_t ws_bytes, int B, int sf, int accumulate, tsr_stream_t stream);
/* nn.Conv2d(Cin -> 1, 3x3, no bias) + ReLU (model/tactileSR_model.py:55-56, :125-126); out is (B,1,H,W) fp32. */
int tsr_tail_fwd(const void* in, int in_ld, int in_bf16, const float* w_oihw, float* out, int B, int H, int W, int Cin,
                 int relu, tsr_stream_t stream);
int tsr_tail_dgrad(const float* dout, const float* out_act, const float* w_oihw, void* din, int din_ld, int din_bf16,
                   int B, int H, int W, int Cin, int relu, tsr_stream_t stream);
/* the same, with the ReLU backward of the layer that produced the tail's input fused in: din = dgrad * [in_act > 0]
   (in_act: that layer's stored activation, in_dtype 1 = bf16 / 2 = fp16) */
int tsr_tail_dgrad_masked(const float* dout, const float* out_act, const float* w_oihw, void* din, int din_ld, int din_bf16,
                          int B, int H, int W, int Cin, int relu, const void* in_act, int in_ld, int in_dtype,
                          tsr_stream_t stream);
size_t tsr_tail_wgrad_workspace(int B, int H, int W, int Cin);
int tsr_tail_wgrad(const void* in, int in_ld, int in_bf16, const float* dout, const float* out_act, float* dw_oihw,
                   void* workspace, size_t ws_bytes, int B, int H, int W, int Cin, int relu, int accumulate,
                   tsr_stream_t stream);

/* ---- BatchNorm / ReLU / layout --------------------------------------------------------------------------------- */
size_t tsr_bn_workspace(long long npix, int C);
/* nn.BatchNorm2d in training mode (model/tactileSR_model.py:38,42,48,169,175,181,187): batch mean / biased variance of
 * y -> scale, shift, saved mean / invstd; running_mean / running_var (momentum, unbiased variance) and
 * num_batches_tracked are updated in place when non-NULL. */
int tsr_bn_train_stats(const void* y, int y_ld, int y_bf16, long long npix, int C, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                       float momentum, float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                       void* workspace, size_t ws_bytes, tsr_stream_t stream);
/* eval mode: scale / shift from the running statistics. */
int tsr_bn_finalize_partials(const float* partial, int part_ld, int nrows, long long npix, int C, const float* gamma,
                             const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                             float momentum, float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                             tsr_stream_t stream);
int tsr_bn_eval_coeffs(int C, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, float* save_mean,
                       float* save_invstd, tsr_stream_t stream);
/* out = [relu](y * scale + shift): BatchNorm2d + nn.ReLU(True), written into a channel slice (replaces torch.cat
 * :74, :81, :200, :203).  out2_bf16 (may be NULL; 16-bit y / out only): a second copy of the result as bf16 with row
 * stride out2_ld -- the weight-gradient operand kept by the "fp16" precision mode. */
int tsr_bn_apply(const void* y, int y_ld, int y_bf16, const float* scale, const float* shift, void* out, int out_ld,
                 int out_bf16, long long npix, int C, int relu, void* out2_bf16, int out2_ld, tsr_stream_t stream);
size_t tsr_bn_backward_workspace(long long npix, int C);
/* autograd backward of BatchNorm2d(+ReLU): da -> dy, dgamma, dbeta (deterministic two-level reductions). */
int tsr_bn_backward(const void* da, int da_ld, const void* y, int y_ld, void* dy, int dy_ld, int act_bf16,
                    const float* scale, const float* shift, const float* save_mean, const float* save_invstd,
                    float* dgamma, float* dbeta, int accumulate, long long npix, int C, int relu, int training,
                    void* workspace, size_t ws_bytes, tsr_stream_t stream);
/* dz = da * [a > 0] for a ReLU fused into a conv epilogue. */
int tsr_relu_backward(const void* da, int da_ld, const void* a, int a_ld, void* dz, int dz_ld, int act_bf16,
                      long long npix, int C, tsr_stream_t stream);
int tsr_copy_channels(const void* x, int x_ld, int x_bf16, void* out, int out_ld, int out_bf16, long long npix, int C,
                      tsr_stream_t stream);
int tsr_nchw_to_nhwc(const float* x, void* out, int out_ld, int out_bf16, int B, int C, int HW, tsr_stream_t stream);
int tsr_nhwc_to_nchw(const void* x, int x_ld, int x_bf16, float* out, int B, int C, int HW, tsr_stream_t stream);

/* ---- loss / optimizer ------------------------------------------------------------------------------------------- */
size_t tsr_mse_hr_workspace(void);
/* train/tactileSR_train.py:44-45,49 fused: HR = bilinear_resize(hr_raw / scale_num, (H,W)); *loss = mean((out-HR)^2);
 * dout = 2 (out - HR) / N * grad_mul (dout may be NULL). */
int tsr_mse_hr_loss(const float* out, const float* hr_raw, float scale_num, int B, int H, int W, int Hin, int Win,
                    float* loss, float* dout, float grad_mul, void* workspace, size_t ws_bytes, tsr_stream_t stream);
/* eval_func (train/tactileSR_train.py:76-94) per-sample metrics with the label preparation fused: sum of squared errors,
 * calculationPSNR and calculationSSIM (utility/tools.py:49-81; psnr_div = the reference's shape[0] * shape[1] of the
 * (1,H,W) slice = H).  out (B,1,H,W), hr_raw (B,1,Hin,Win) -> sqerr / psnr / ssim (B each). */
int tsr_eval_metrics(const float* out, const float* hr_raw, float scale_num, int B, int H, int W, int Hin, int Win,
                     float max_value, float psnr_div, float c1, float c2, float* sqerr, float* psnr, float* ssim,
                     tsr_stream_t stream);
/* torch.optim.Adam step (coupled L2 weight decay, no amsgrad; train/tactileSR_train.py:212, tPSFNet_train.py:201)
 * over a flat fp32 buffer of n elements; `step` is the 1-based step count, lr a host scalar read every call. */
int tsr_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, long long step, float grad_scale, tsr_stream_t stream);

/* the same step with the step-dependent scalars in device memory, hyper = {lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)}: a
 * captured CUDA graph can replay the launch while the host advances t and the learning-rate schedule. */
int tsr_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper, float beta1,
                      float beta2, float eps, float weight_decay, float grad_scale, tsr_stream_t stream);

/* ---- tPSFNet ----------------------------------------------------------------------------------------------------- */
/* C[m][n] = act(sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] + bias[n]); act 0 none, 1 relu, 2 softplus. */
int tsr_sgemm_strided(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                      float* C, long long ldc, int M, int N, int K, const float* bias, int act, int accumulate,
                      tsr_stream_t stream);
/* nn.Linear (+ReLU / Softplus) of MLP_layer (model/tPSFNet.py:26-36) and its backward. */
int tsr_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int N, int K, int act,
                   tsr_stream_t stream);
size_t tsr_linear_bwd_workspace(int M, int N, int K);
int tsr_linear_bwd(const float* dy, const float* out, const float* x, const float* w, float* dpre, float* dw, float* db,
                   float* dx, int M, int N, int K, int act, int accumulate, void* workspace, size_t ws_bytes,
                   tsr_stream_t stream);
/* The per-sample loop of tPSFNet.forward (model/tPSFNet.py:118-125): tactilePSF :78-83, depth2tactile :85-100 (dense
 * 99x99 correlation + second-max fill) and degradation_process :129-141, fused, one CTA per sample.
 * alphaBeta (B,3), depth (B,100,100) -> HR (B,100,100), LRd (B,16), psf (B,99,99; may be NULL). */
int tsr_psf_forward(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                    tsr_stream_t stream);
/* the implementations behind tsr_psf_forward: tcgen05 with fp16 hi/lo split operands (fp32-accurate; default, mode 0),
 * FFMA (mode 1) and tcgen05 with ONE fp16 pass (~3e-4; the 16-bit tensor-core precision modes, mode 2);
 * tsr_set_psf_mode(0 | 1 | 2) selects which one tsr_psf_forward runs.  The tcgen05 forwards can leave per-row statistics
 * of HR in `aux` (B x tsr_psf_aux_floats() floats, 16-byte aligned, may be NULL) for tsr_psf_backward_tc[_f16], the tcgen05
 * backward of the training case (gradient through LR_degrade only, train/tPSFNet_train.py:186-189). */
size_t tsr_psf_aux_floats(void);
int tsr_psf_forward_tc(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, float* aux, int B,
                       tsr_stream_t stream);
int tsr_psf_backward_tc(const float* alphaBeta, const float* depth, const float* aux, const float* dLRd,
                        float* dalphaBeta, int B, tsr_stream_t stream);
int tsr_psf_forward_tc_f16(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, float* aux, int B,
                           tsr_stream_t stream);
int tsr_psf_backward_tc_f16(const float* alphaBeta, const float* depth, const float* aux, const float* dLRd,
                            float* dalphaBeta, int B, tsr_stream_t stream);
int tsr_psf_forward_ffma(const float* alphaBeta, const float* depth, float* HR, float* LRd, float* psf, int B,
                         tsr_stream_t stream);
void tsr_set_psf_mode(int mode);
int tsr_get_psf_mode(void);
/* its backward: d alphaBeta (B,3) from dLRd (B,16), dHR (B,100,100), dpsf (B,99,99) -- each may be NULL (= 0). */
int tsr_psf_backward(const float* alphaBeta, const float* depth, const float* HR, const float* dLRd, const float* dHR,
                     const float* dpsf, float* dalphaBeta, int B, tsr_stream_t stream);

/* ---- tensor-core convolutions (tcgen05 / TMEM / TMA, bf16 operands, fp32 accumulate) ------------------------------ */
/* OIHW fp32 -> bf16 [ci chunk][tap][Cout][64] tiles in the SWIZZLE_128B shared-memory image (forward) and the flipped /
 * transposed set for the data gradient.  Cin, Cout % 64 == 0. */
int tsr_pack_conv_weight_bf16(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int KS,
                              tsr_stream_t stream);
/* every conv weight of a model in one launch (they all change at each optimizer step): a device-memory table of */
typedef struct TsrPackDesc {
  const float* w;   /* OIHW fp32 */
  void* wf;         /* forward image or NULL */
  void* wd;         /* data-gradient image or NULL */
  int Cout, Cin, KS;
  int dt_f, dt_d;   /* storage codes of wf / wd: 1 = bf16, 2 = fp16 */
  int mode;         /* 0: wf = standard forward image; 1 / 2: wf = a dual-branch image (tsr_pack_conv_weight_dual) and this
                       weight is its 3x3 / 5x5 branch */
} TsrPackDesc;
int tsr_pack_conv_weights_multi(const void* table_dev, int n, long long max_elems, tsr_stream_t stream);
/* same packing with fp16 elements: forward weights of the "fp16" precision mode (fp16 activations, bf16 gradients) */
int tsr_pack_conv_weight_f16(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int KS,
                             tsr_stream_t stream);
/* inference: fold an eval-mode BatchNorm (scale / shift of tsr_bn_eval_coeffs) that follows the convolution into the packed
   forward weights (dtype 1 = bf16, 2 = fp16) and into bias_out[co] = scale[co] * bias[co] + shift[co] (bias may be NULL) */
int tsr_pack_conv_weight_folded(const float* w_oihw, const float* bias, const float* scale, const float* shift, void* w_fwd,
                                float* bias_out, int Cout, int Cin, int KS, int dtype, tsr_stream_t stream);
size_t tsr_conv2d_tc_workspace(int B, int H, int W, int Cin, int Cout, int KS);
/* bf16 NHWC convolution, same semantics as tsr_conv2d_f32 (bias fp32, residual / out bf16).  W % 8 == 0.
   flags bit 1 (value 2): in / weights / residual / out are fp16 instead of bf16 (same kernels, kind::f16 format field). */
int tsr_conv2d_tc(const void* in, int in_ld, const void* w_packed, const float* bias, const void* residual, int res_ld,
                  void* out, int out_ld, int B, int H, int W, int Cin, int Cout, int KS, int flags, void* workspace,
                  size_t ws_bytes, float* bn_partial, void* out2_bf16, int out2_ld, tsr_stream_t stream);
/* out2_bf16 (may be NULL): a second copy of the result as bf16 with row stride out2_ld (weight-gradient operand kept by the
   "fp16" precision mode). */
/* bn_partial (may be NULL): [tsr_conv2d_tc_stat_rows()][2][Cout] floats receiving partial sums / sums of squares of the
   stored output from the epilogue (batch statistics of the BatchNorm that follows, nn.BatchNorm2d at
   tactileSR_model.py:42,48,169,...); finish them with tsr_bn_finalize_partials. */
int tsr_conv2d_tc_stat_rows(void);
size_t tsr_conv2d_wgrad_tc_workspace(int B, int H, int W, int Cin, int Cout, int KS);
/* weight gradient (fp32 OIHW) of bf16 activations / gradients; bit-deterministic.  H, W % 8 == 0, Cout = 64 or a multiple of 128
   (128 output channels per launch).
   (The "fp16" precision mode hands this a bf16 copy of the conv input: tcgen05 kind::f16 rejects fp16 x bf16.) */
int tsr_conv2d_wgrad_tc(const void* in, int in_ld, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
                        size_t ws_bytes, int B, int H, int W, int Cin, int Cout, int KS, int accumulate,
                        tsr_stream_t stream);

/* ---- tensor-core convolution, generation 2 (csrc/conv_tc2.cu) -------------------------------------------------
 * One launch = out = epilogue( sum over sources s, taps, channels of in_s[pix + shift] * w_s ), the sources accumulating
 * into one accumulator:
 *   nsrc = 1                the forward (or data gradient) of one nn.Conv2d, as tsr_conv2d_tc;
 *   nsrc = 2                K-concatenation: the data gradient of the two convolutions that read the same tensor
 *                           (MSRB conv_3_x / conv_5_x, model/tactileSR_model.py:198-199, 201-202) in one pass;
 *   nsrc = 1, dual_fwd = 1  the 3x3 and the 5x5 convolution (64 output channels each) of one input from one halo tile:
 *                           out channels [0, 64) = conv3, [64, 128) = conv5; src[0].KS = 5 and src[0].w_packed is the
 *                           image of tsr_pack_conv_weight_dual.
 * Epilogue, in this order: + bias, + residual, [mask], [ReLU], round to the output type (fp16 with TSR_TC2_F16, else
 * bf16), store (+ bf16 copy out2), statistics.
 *   TSR_TC2_MASK      aux = a saved activation (fp16 with TSR_TC2_AUX_F16, else bf16): the value is zeroed where
 *                     aux <= 0 -- the ReLU backward of the layer that produced `aux` (replaces tsr_relu_backward);
 *   TSR_TC2_BNB       aux = y, the input of a BatchNorm(+ReLU): with TSR_TC2_BNB_RELU the value is zeroed where
 *                     aux_scale*y + aux_shift <= 0, and stat receives per-row partial sums of (g, g*y) over the stored g
 *                     -- level 1 of the BatchNorm backward (finish with tsr_bn_bwd_finalize_partials, then
 *                     tsr_bn_backward_apply with relu = 0);
 *   otherwise         stat (may be NULL) receives partial sums of (o, o*o) over the stored output o -- the batch
 *                     statistics of the BatchNorm that follows (finish with tsr_bn_finalize_partials).
 * stat: [tsr_conv2d_tc2_stat_rows()][2][stat_ld] floats; cleared by the call unless TSR_TC2_STAT_PRECLEARED (required when
 * `stat` points into a channel slice of a wider table).  Bit-deterministic. */
#define TSR_TC2_RELU 1
#define TSR_TC2_F16 2
#define TSR_TC2_MASK 8
#define TSR_TC2_BNB 16
#define TSR_TC2_BNB_RELU 32
#define TSR_TC2_AUX_F16 64
#define TSR_TC2_STAT_PRECLEARED 128
typedef struct TsrConvSrc {
  const void* in;        /* NHWC, Cin channels from this pointer */
  const void* w_packed;  /* tsr_pack_conv_weight_bf16 / _f16 image (forward or data-gradient), or the dual image */
  int in_ld, Cin, KS, pad_;
} TsrConvSrc;
typedef struct TsrConvTc2 {
  TsrConvSrc src[2];
  const float* bias;       /* [Cout] or NULL */
  const void* residual;    /* [pix][res_ld], output type, or NULL */
  void* out;               /* [pix][out_ld] */
  void* out2_bf16;         /* optional bf16 copy of an fp16 output */
  float* stat;
  const void* aux;
  const float* aux_scale;
  const float* aux_shift;
  int nsrc, dual_fwd;
  int res_ld, out_ld, out2_ld, stat_ld, aux_ld;
  int B, H, W, Cout, flags;
} TsrConvTc2;
int tsr_conv2d_tc2(const void* args /* const TsrConvTc2*, host memory */, tsr_stream_t stream);
int tsr_conv2d_tc2_stat_rows(void);
/* diagnostics: 8 device counters filled by CTA 0 of the following tsr_conv2d_tc2 launches with the cycles its warp roles
   spent waiting (see csrc/conv_tc2.cu); NULL switches it off */
void tsr_conv2d_tc2_debug(unsigned long long* counters_dev);
/* w3 (64, Cin, 3, 3), w5 (64, Cin, 5, 5) fp32 OIHW -> dual-branch forward image (dtype 1 = bf16, 2 = fp16) of
   tsr_pack_conv_weight_dual_elems(Cin) elements */
size_t tsr_pack_conv_weight_dual_elems(int Cin);
int tsr_pack_conv_weight_dual(const float* w3, const float* w5, void* out, int Cin, int dtype, tsr_stream_t stream);
/* BatchNorm backward split for the fused data-gradient epilogue: level 2 from a (sum g, sum g*y) table, and level 3
   (dy = scale * (g - c1 - xhat * c2)) on its own.  c1, c2: [C] floats.  nn.BatchNorm2d backward, cpu/trainer.py:353. */
int tsr_bn_bwd_finalize_partials(const float* partial, int part_ld, int nrows, long long npix, int C, const float* save_mean,
                                 const float* save_invstd, float* dgamma, float* dbeta, int accumulate, float* c1, float* c2,
                                 int training, tsr_stream_t stream);
int tsr_bn_backward_apply(const void* da, int da_ld, const void* y, int y_ld, void* dy, int dy_ld, int act_bf16,
                          const float* scale, const float* shift, const float* save_mean, const float* save_invstd,
                          const float* c1, const float* c2, long long npix, int C, int relu, tsr_stream_t stream);

/* tsr_conv2d_wgrad_tc with the storage type of `in` given (1 = bf16, 2 = fp16): an fp16 input -- the forward activations of the
   "fp16" precision mode -- is converted to bf16 in shared memory by otherwise idle warps before the MMAs read it, so no
   bf16 copy of the conv input has to exist in HBM.  dout stays bf16. */
int tsr_conv2d_wgrad_tc_x(const void* in, int in_ld, int in_dtype, const void* dout, int dout_ld, float* dw_oihw, void* workspace,
                          size_t ws_bytes, int B, int H, int W, int Cin, int Cout, int KS, int accumulate,
                          tsr_stream_t stream);
/* experiment switches of the tensor-core kernels (see csrc/conv_tc.cu); 0 = production configuration. */
void tsr_set_tc_desc_mode(int mode);
int tsr_get_tc_desc_mode(void);

#ifdef __cplusplus
}
#endif
#endif /* TACTILESR_B200_H */
