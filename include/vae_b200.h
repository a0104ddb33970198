/*
 * vae_b200.h -- C ABI of the B200-native VanillaVAE train-step kernels.
 *
 * The reference (finlaymiller/torch-vae) has no FFI: its hot path is the Python
 * surface of midi_autoencoder/models.py executed by ATen.  Each entry point
 * below replaces the ATen op(s) that a reference line dispatches to; the
 * "replaces" note on every prototype cites that line (file:line relative to
 * /root/reference/midi_autoencoder/).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *    it says "host"; `stream` is a cudaStream_t passed as void*.
 *  - activations are NHWC ("channels last"), dtype code VAE_F32 or VAE_BF16;
 *    parameters, BatchNorm vectors, latents and gradients of parameters are
 *    always fp32 in the reference's own layouts (OIHW / [out,in]).
 *  - return value: 0 on success, <0 on error (VAE_ERR_*); vae_last_error()
 *    gives a host string.  No C++ exception crosses this boundary; the Python
 *    host raises RuntimeError, the reference's own error convention.
 *  - nothing here allocates device memory: workspaces are caller-provided
 *    (the host side takes them from torch's caching allocator).
 */
#ifndef VAE_B200_H_
#define VAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAE_ABI_VERSION 1

#define VAE_F32 0
#define VAE_BF16 1

#define VAE_OK 0
#define VAE_ERR_ARG (-1)
#define VAE_ERR_CUDA (-2)
#define VAE_ERR_UNSUPPORTED (-3)

typedef void* vae_stream_t;

/* Leave `n` SMs out of the grids of the persistent kernels, for NCCL kernels that run concurrently in data-parallel runs
 * with large gradient buckets (the persistent kernels otherwise occupy every SM completely); returns the previous value. */
int vae_set_reserved_sms(int n);
/* Scratch arena the caller keeps zeroed (current device): buffers inside it that a kernel accumulates into are not cleared
 * by the library; base = NULL removes the registration.  Host-side only, no kernel launch. */
int vae_set_zero_arena(void* base, long long bytes);
int vae_abi_version(void);
const char* vae_last_error(void);
/* 1 if the library was built with the tcgen05/TMEM (UMMA) conv kernels. */
int vae_has_umma(void);

/* ------------------------------------------------------------------------
 * Convolution engine (3x3, pad 1, stride 1 or 2).
 *
 * One gather-GEMM kernel family serves Conv2d forward, ConvTranspose2d
 * forward (= data gradient of the equivalent Conv2d) and both data gradients;
 * one split-K kernel family serves every weight gradient.  In "equivalent
 * conv" terms a layer maps X[B,Hi,Wi,Ci] -> Y[B,Hi/s,Wi/s,Co] with weight
 * W[Co,Ci,3,3]; ConvTranspose2d(cin,cout,k3,s2,p1,op1) (models.py:66-68,77)
 * is the same thing read backwards (Co=cin, Ci=cout, same memory layout).
 * ---------------------------------------------------------------------- */

/* Repack an fp32 OIHW weight into the two K-major operand layouts:
 *   wp[Co][3][3][Ci]  ("fprop" operand)   and   wt[Ci][3][3][Co] ("dgrad" operand)
 * in dtype `dtype`.  Either destination may be NULL.
 * replaces: the weight layout cuDNN/MKLDNN pick internally for models.py:45,66,77,80 */
int vae_pack_conv_weight(const float* w_oihw, void* wp, void* wt, int Co, int Ci, int dtype,
                         vae_stream_t stream);

/* y = conv(x, w) + bias; optionally accumulates per-channel sum / sum-of-squares
 * of y into stats[2*Co] (double; zeroed by this call) for train-mode BatchNorm.
 * dt_in may differ from dt_out (the fp32 piano-roll feeds a bf16 network).
 * replaces: aten::convolution at models.py:45 (encoder) and :80 (final conv) */
int vae_conv_fprop(const void* x, const void* wp, const float* bias, void* y, double* stats,
                   int B, int Hi, int Wi, int Ci, int Co, int stride, int dt_in, int dt_out,
                   vae_stream_t stream);

/* dx = conv_data_grad(dy, w) (+ bias, + stats over dx when used as the
 * ConvTranspose2d forward).  dy[B,Hi/s,Wi/s,Co], wt[Ci][3][3][Co], dx[B,Hi,Wi,Ci].
 * replaces: aten::convolution (transposed) at models.py:66-68,77 and the
 * data-gradient half of aten::convolution_backward (train.py:650) */
int vae_conv_dgrad(const void* dy, const void* wt, const float* bias, void* dx, double* stats,
                   int B, int Hi, int Wi, int Ci, int Co, int stride, int dtype,
                   vae_stream_t stream);

/* dw[Co][Ci][3][3] (fp32, OIHW, overwritten) = sum over pixels of dy (x) patches(x).
 * replaces: the weight-gradient half of aten::convolution_backward (train.py:650) */
int vae_conv_wgrad(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Ci,
                   int Co, int stride, int dt_x, int dt_dy, vae_stream_t stream);

/* ---- tcgen05 / TMEM path (bf16; Cin_gemm % 32 == 0, 16 <= N <= 256, N % 16 == 0) ----
 * mode 0 = conv forward, mode 1 = data gradient (== ConvTranspose2d forward).
 * Weights are pre-packed once per optimizer step into the exact 128B-swizzled
 * shared-memory images the kernel's bulk-TMA copies fetch. */
int vae_set_umma(int enable); /* returns the previous setting; tests use it to cross-check */
int vae_umma_set_trace(void* buf); /* tuning aid: 6*512 int64 clock64 samples from CTA 0; NULL = off */
long long vae_umma_packed_elems(int Ci, int Co, int stride, int mode); /* 0: shape not served */
int vae_umma_pack_weight(const float* w_oihw, void* out, int Ci, int Co, int stride, int mode,
                         vae_stream_t stream);
/* in/out NHWC bf16.  Hi, Wi: HIGH-resolution extents (as in vae_conv_fprop/dgrad).
 * pro_scale/pro_mean/pro_beta (nullable, fp32[C of `in`]) apply the PREVIOUS layer's
 * BatchNorm+LeakyReLU, leaky_relu((x-mean)*scale+beta), to the operand while it is
 * gathered (padding stays zero).
 * replaces: aten::convolution / transposed convolution, models.py:45,66-68,77 */
int vae_conv_umma(const void* in, const void* wpk, const float* bias, void* out, double* stats,
                  const float* pro_scale, const float* pro_mean, const float* pro_beta, float slope,
                  int B, int Hi, int Wi, int Ci, int Co, int stride, int mode, vae_stream_t stream);

/* TMA-fed variant of vae_conv_umma: the A operand of every 128-pixel tile is staged by
 * cp.async.bulk.tensor (one strided box of the NHWC tensor per filter tap, zero-filled
 * outside the image) instead of cp.async threads.  Serves layers whose extents decompose
 * into whole output rows per tile and Cin == 32 or Cin % 64 == 0; no fused prologue. */
/* 1 / 0: use the tensor-core kernel for the single-input-channel edge layer inside vae_conv_fprop (Ci = 1, Co = 32 / 64,
 * stride 2, fp32 in, bf16 out); returns the previous setting. */
int vae_set_edge_tc(int enable);
/* Bit 0: serve transposed stride-2 convs with 64 output channels (mode 1 of vae_conv_umma_tma) by the kernel that fuses the
 * four output-parity phases (conv_umma_tmat.cu); bit 1: also 128 output channels (measured slower; for tests).  Also selects the layout vae_umma_tma_pack_weight produces
 * for those layers: flip it BEFORE packing.  Returns the previous setting. */
int vae_set_umma_tmat(int enable);
int vae_set_umma_tma(int enable);
int vae_umma_tma_eligible(int B, int Hi, int Wi, int Ci, int Co, int stride, int mode);
long long vae_umma_tma_packed_elems(int Ci, int Co, int stride, int mode);
int vae_umma_tma_pack_weight(const float* w_oihw, void* out, int Ci, int Co, int stride, int mode,
                             vae_stream_t stream);
int vae_conv_umma_tma(const void* in, const void* wpk, const float* bias, void* out, double* stats,
                      int B, int Hi, int Wi, int Ci, int Co, int stride, int mode,
                      vae_stream_t stream);

/* Weight gradient on the tcgen05 path (both operands MN-major, split over pixels,
 * fp32 red.global.add into dw).  x, dy NHWC bf16; dw [Co][Ci][3][3] fp32, overwritten.
 * replaces: the weight-gradient half of aten::convolution_backward (train.py:650) */
int vae_set_umma_wgrad(int enable);
int vae_umma_wgrad_supported(int Ci, int Co);
int vae_conv_wgrad_umma(const void* x, const void* dy, float* dw, int B, int Hi, int Wi, int Ci,
                        int Co, int stride, vae_stream_t stream);

/* Resident-weight TMA/tcgen05 conv for the narrow layers (conv_umma_rw.cu): all nine weight
 * images stay in shared memory, stages carry activations only; mode 1 computes the four
 * output-parity phases of ConvTranspose2d(k3,s2,p1,op1) per tile from the same four boxes.
 * Serves mode 0: Ci == 32, Co in {32,64}; mode 1: Ci == 32 (outputs), Co in {32,64}, stride 2.
 * Same contract as vae_conv_umma_tma; wpk (9*Ci*Co bf16) from vae_umma_rw_pack_weight.
 * replaces: aten::convolution (models.py:45,66-68,77) and the data-gradient half of
 * aten::convolution_backward (train.py:650) on those layers */
int vae_set_umma_rw(int enable);
int vae_umma_rw_eligible(int B, int Hi, int Wi, int Ci, int Co, int stride, int mode);
long long vae_umma_rw_packed_elems(int Ci, int Co);
int vae_umma_rw_pack_weight(const float* w_oihw, void* out, int Ci, int Co, int mode,
                            vae_stream_t stream);
int vae_conv_umma_rw(const void* in, const void* wpk, const float* bias, void* out, double* stats,
                     int B, int Hi, int Wi, int Ci, int Co, int stride, int mode,
                     vae_stream_t stream);

/* The same two kernels for INFERENCE (model.eval() under torch.no_grad(): models.py:250-263, evaluation.py:42-65):
 * eval-mode BatchNorm (running statistics: mean, scale = gamma / sqrt(running_var + eps), beta) + LeakyReLU applied in
 * the epilogue, out = lrelu((acc + bias - mean) * scale + beta): the block writes its ACTIVATION, no statistics, no
 * separate normalisation pass.
 * replaces: aten::convolution + aten::native_batch_norm (eval) + aten::leaky_relu of one block (models.py:45-47,66-70) */
int vae_conv_umma_rw_act(const void* in, const void* wpk, const float* bias, void* out, const float* mean,
                         const float* scale, const float* beta, float slope, int B, int Hi, int Wi, int Ci, int Co,
                         int stride, int mode, vae_stream_t stream);
int vae_conv_umma_tma_act(const void* in, const void* wpk, const float* bias, void* out, const float* mean,
                          const float* scale, const float* beta, float slope, int B, int Hi, int Wi, int Ci, int Co,
                          int stride, int mode, vae_stream_t stream);

/* The same two kernels producing the gradient w.r.t. an activation act = lrelu(bn(y)) of the layer below
 * (models.py:45-47,66-70) with the first half of that BatchNorm's backward fused into the epilogue: `y` is the
 * stored pre-normalisation tensor of the layer below (same shape as `out`), mean / scale (= gamma*rstd) / beta / rstd its
 * [channels] fp32 vectors.  `out` receives du = d_act * LeakyReLU'(u) and dstats[0..C) = sum du,
 * dstats[C..2C) = sum du * xhat (fp64) -- what vae_bn_lrelu_bwd_reduce computes in a separate pass over
 * (d_act, y); follow with vae_bn_lrelu_bwd_apply(training | 2).
 * replaces: aten::convolution_backward (data gradient) + aten::leaky_relu_backward + the reductions of
 * aten::native_batch_norm_backward (train.py:650) */
int vae_conv_umma_rw_bnbwd(const void* in, const void* wpk, void* out, const void* y, const float* mean,
                           const float* scale, const float* beta, const float* rstd, double* dstats, float slope,
                           int B, int Hi, int Wi, int Ci, int Co, int stride, int mode, vae_stream_t stream);
int vae_conv_umma_tma_bnbwd(const void* in, const void* wpk, void* out, const void* y, const float* mean,
                            const float* scale, const float* beta, const float* rstd, double* dstats, float slope,
                            int B, int Hi, int Wi, int Ci, int Co, int stride, int mode, vae_stream_t stream);

/* TMA-fed weight gradient (conv_wgrad_tma.cu): operands staged MN-major straight from NHWC
 * boxes; ws is a caller-provided fp32 workspace of 9*Co*Ci elements; dw overwritten.
 * replaces: the weight-gradient half of aten::convolution_backward (train.py:650) */
int vae_set_wgrad_tma(int enable);
int vae_wgrad_tma_eligible(int B, int Hi, int Wi, int Ci, int Co, int stride);
int vae_conv_wgrad_tma(const void* x, const void* dy, float* dw, float* ws, int B, int Hi, int Wi,
                       int Ci, int Co, int stride, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * BatchNorm2d (train and eval) + LeakyReLU.   models.py:46-47,69-70,78-79
 * ---------------------------------------------------------------------- */

/* From stats[2C] (sum, sumsq over `count` values per channel) produce
 * mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, and update the
 * running buffers (momentum, unbiased variance) when `training`.  With
 * training == 0 the running buffers are used instead of stats (eval mode).
 * replaces: the statistics half of aten::native_batch_norm */
int vae_bn_finalize(const double* stats, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float* scale, float* shift,
                    float* mean, float* rstd, int C, long long count, float momentum, float eps,
                    int training, vae_stream_t stream);

/* act = leaky_relu((y - mean[c]) * scale[c] + beta[c])  (ATen's operation order, so the
 * LeakyReLU branch agrees with the reference down to |u| ~ 1e-7).
 * replaces: the normalise half of aten::native_batch_norm + aten::leaky_relu */
int vae_bn_apply_lrelu(const void* y, const float* scale, const float* mean, const float* beta,
                       void* act, long long n_pix, int C, float slope, int dtype,
                       vae_stream_t stream);

/* vae_bn_finalize + vae_bn_apply_lrelu in one launch (C % 8 == 0): vec_out[4][C] receives
 * (scale, shift, mean, rstd); num_batches_tracked (int64 device scalar, may be NULL) is
 * incremented when training.
 * replaces: aten::native_batch_norm + aten::leaky_relu (models.py:46-47,69-70) */
int vae_bn_finalize_apply_lrelu(const void* y, const double* stats, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, long long* num_batches_tracked,
                                float* vec_out, void* act, long long n_pix, int C, long long count,
                                float momentum, float eps, int training, float slope, int dtype,
                                vae_stream_t stream);

/* Pass 1 of the backward: du = dact * lrelu'((y-mean)*scale+beta);
 * dstats[0:C] += sum du, dstats[C:2C] += sum du * xhat (double; zeroed by this call).
 * replaces: aten::leaky_relu_backward + the reductions of native_batch_norm_backward */
int vae_bn_lrelu_bwd_reduce(const void* dact, const void* y, const float* scale,
                            const float* beta, const float* mean, const float* rstd,
                            double* dstats, long long n_pix, int C, float slope, int dtype,
                            vae_stream_t stream);

/* Pass 2: dy = gamma*rstd*(du - dbeta/M - xhat*dgamma/M) (training) or
 * dy = du*scale (eval); writes dgamma[C], dbeta[C] and dbias[C] = sum dy (the
 * gradient of the conv bias that feeds this BatchNorm).  `training`: bit 0 = train-mode
 * statistics, bit 1 = `dact` already holds du (written by vae_conv_umma_*_bnbwd, whose
 * epilogue also produced dstats, so pass 1 is skipped).
 * replaces: aten::native_batch_norm_backward */
int vae_bn_lrelu_bwd_apply(const void* dact, const void* y, const float* scale,
                           const float* beta, const float* mean, const float* rstd,
                           const float* gamma, const double* dstats, void* dy, float* dgamma,
                           float* dbeta, float* dbias, long long n_pix, int C, float slope,
                           int training, int dtype, vae_stream_t stream);

/* Both passes in ONE cooperative launch for tensors that stay in the L2 between them (d_act + y <= 96 MB): reduce,
 * grid-wide arrival counter, apply on the same pixels.  dstats: [2C + 1] doubles, zeroed by the call.  Returns
 * VAE_ERR_UNSUPPORTED (-3) when the shape should take the two-pass path.
 * replaces: aten::leaky_relu_backward + aten::native_batch_norm_backward (train.py:650) */
int vae_bn_lrelu_bwd_fused(const void* dact, const void* y, const float* scale, const float* beta,
                           const float* mean, const float* rstd, const float* gamma, double* dstats, void* dy,
                           float* dgamma, float* dbeta, float* dbias, long long n_pix, int C, float slope,
                           int training, int dtype, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused output block of the decoder (models.py:76-82, run at :171):
 *   y = ConvT(..) [B,H,W,C] -> BatchNorm2d -> LeakyReLU -> Conv2d(C->1,k3,s1,p1) -> Sigmoid.
 * y is the largest tensor of the network; these three entry points read it once
 * each and never store the normalised activation or its gradient.
 * mean/scale/beta/rstd are the BatchNorm vectors of vae_bn_finalize
 * (u = (y-mean)*scale+beta); w_last is the fp32 OIHW weight [1,C,3,3] as torch
 * holds it (no repacking); C in {8,16,32,64}, W % 4 == 0.
 * ---------------------------------------------------------------------- */

/* 1 when the three entry points below serve this shape. */
int vae_final_block_supported(int B, int H, int W, int C);

/* logits[B,H,W] (dtype) = conv(lrelu(bn(y)), w_last) + b_last; probs[B,H,W] (fp32,
 * may be NULL) = sigmoid(logits).
 * replaces: aten::native_batch_norm (apply) + leaky_relu + convolution + sigmoid at
 * models.py:78-81 */
int vae_final_conv_fwd(const void* y, const float* mean, const float* scale, const float* beta,
                       const float* w_last, const float* b_last, void* logits, float* probs, int B,
                       int H, int W, int C, float slope, int dtype, vae_stream_t stream);

/* Backward pass 1.  With da = conv_data_grad(dlogits, w_last) and du = da*lrelu'(u):
 *   dstats[0:C] = sum du, dstats[C:2C] = sum du*xhat, dstats[2C] = sum dlogits
 *   (double[2C+1], zeroed by this call);  dw_last[C*9] = weight gradient of the final
 *   conv (fp32 OIHW, zeroed by this call).
 * replaces: convolution_backward (models.py:80) + leaky_relu_backward + the reductions
 * of native_batch_norm_backward (models.py:78) inside train.py:650 */
int vae_final_bwd_reduce(const void* dlogits, const void* y, const float* mean, const float* scale,
                         const float* beta, const float* rstd, const float* w_last, double* dstats,
                         float* dw_last, int B, int H, int W, int C, float slope, int dtype,
                         vae_stream_t stream);

/* Backward pass 2: dy[B,H,W,C] = gamma*rstd*(du - dstats[c]/M - xhat*dstats[C+c]/M)
 * (training) or du*gamma*rstd (eval); also writes dgamma[C], dbeta[C], dbias[C] = sum dy
 * (bias of the ConvT feeding the BatchNorm) and db_last[1] (bias of the final conv).
 * replaces: native_batch_norm_backward (models.py:78) inside train.py:650 */
int vae_final_bwd_apply(const void* dlogits, const void* y, const float* mean, const float* scale,
                        const float* beta, const float* rstd, const float* gamma, const float* w_last,
                        const double* dstats, void* dy, float* dgamma, float* dbeta, float* dbias,
                        float* db_last, int B, int H, int W, int C, float slope, int training,
                        int dtype, vae_stream_t stream);

/* tcgen05 / TMEM versions of the three output-block kernels (final_block_tc.cu) for bf16,
 * C == 32, W == 128: per pixel the C->1 conv is P[p][tap] = act[p][:] . w[tap][:] (a GEMM with
 * M = pixels, N = 9, K = C) followed by a 9-point shifted sum; its data gradient is a GEMM with
 * K = 9.  Same contracts as the entry points above. */
int vae_set_final_tc(int enable);
int vae_final_tc_supported(int B, int H, int W, int C, int dtype);
int vae_final_conv_fwd_tc(const void* y, const float* mean, const float* scale, const float* beta,
                          const float* w_last, const float* b_last, void* logits, float* probs,
                          int B, int H, int W, int C, float slope, vae_stream_t stream);
int vae_final_bwd_reduce_tc(const void* dlogits, const void* y, const float* mean, const float* scale,
                            const float* beta, const float* rstd, const float* w_last, double* dstats,
                            float* dw_last, int B, int H, int W, int C, float slope,
                            vae_stream_t stream);
int vae_final_bwd_apply_tc(const void* dlogits, const void* y, const float* mean, const float* scale,
                           const float* beta, const float* rstd, const float* gamma, const float* w_last,
                           const double* dstats, void* dy, float* dgamma, float* dbeta, float* dbias,
                           float* db_last, int B, int H, int W, int C, float slope, int training,
                           vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused AdamW: one launch for all tensors of the reference's two
 * parameter groups; state layout of torch.optim.AdamW(capturable=True).
 * replaces: optimizer.step() at train.py:656 (torch.optim.AdamW built at :210-228)
 * ---------------------------------------------------------------------- */
typedef struct {
  void* p;          /* parameter, fp32, n elements (device) */
  const void* g;    /* gradient */
  void* m;          /* exp_avg */
  void* v;          /* exp_avg_sq */
  float* step;      /* device scalar, float */
  long long n;
  float lr;
  float wd;
} vae_adamw_tensor;
int vae_adamw_step(const vae_adamw_tensor* tensors /* host array */, int count, double beta1, double beta2,
                   double eps, vae_stream_t stream);

/* The same step with the hyper-parameters in DEVICE memory, so that a captured CUDA graph follows a
 * per-batch schedule: hyper[8*g .. 8*g+5] = {lr, beta1, beta2, weight_decay, 1-beta1, 1-beta2} (fp32, rows of 8) of group g,
 * group_of[i] (host array) = group of tensor i; the lr / wd fields of the records are ignored.
 * replaces: optimizer.step() (train.py:656) under the per-batch scheduler of train.py:232-238,659 */
int vae_adamw_step_dev(const vae_adamw_tensor* tensors /* host array */, const int* group_of /* host array */,
                       int count, const float* hyper, int n_groups, double eps, vae_stream_t stream);

/* One scheduler.step() of torch.optim.lr_scheduler.OneCycleLR (two phases) on the device:
 * counter[0] (int64 = last_epoch) += 1, then hyper[8*g] = lr and, with cycle_momentum, hyper[8*g+1] = beta1, hyper[8*g+4] = 1 - beta1.
 * per_group (host array) = n_groups (<= 8) records {initial_lr, max_lr, min_lr, base_momentum, max_momentum}
 * of doubles; anneal 0 = cos, 1 = linear.  Past total_steps (where torch raises) the schedule is clamped.
 * replaces: scheduler.step() (train.py:659) of the OneCycleLR built at train.py:232-238 */
int vae_onecycle_step(float* hyper, long long* counter, const double* per_group /* host array */, int n_groups,
                      long long total_steps, double pct_start, int anneal, int cycle_momentum, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Skinny GEMMs of the latent heads for L <= 16 (models.py:55-59,137,141,162 and their
 * gradients inside train.py:650).  fp32 master weights are used in place:
 * W(n, k) = W[n*wsn + k*wsk] covers both [L,F] (fc_mu / fc_var) and [F,L]
 * (decoder_input) layouts.  K (= F) must be a multiple of 8.
 * ---------------------------------------------------------------------- */
int vae_heads_supported(int F, int L);

/* out[M][N] fp32 = A[M][K] (dtype dtA) . W + bias.
 * replaces: aten::addmm at models.py:137,141 and dL/dz of decoder_input */
int vae_skinny_reduce(const void* A, const float* W, const float* bias, float* out, int M, int N,
                      int K, long long wsn, long long wsk, int dtA, vae_stream_t stream);
/* Two heads that share their input in ONE pass over it (fc_mu and fc_var, models.py:137,141): out0 = A . W0 + bias0,
 * out1 = A . W1 + bias1, same shapes and strides as vae_skinny_reduce.
 * replaces: the two aten::addmm of fc_mu / fc_var */
int vae_skinny_reduce_pair(const void* A, const float* W0, const float* bias0, float* out0, const float* W1,
                           const float* bias1, float* out1, int M, int N, int K, long long wsn, long long wsk,
                           int dtA, vae_stream_t stream);

/* G[n*gsn + k*gsk] += sum_m D[m][n]*A[m][k]; colsum[k] += sum_m A[m][k] (may be NULL).
 * G and colsum must be zeroed by the caller.
 * replaces: weight / bias gradients of the three Linear layers */
int vae_skinny_wgrad(const float* D, const void* A, float* G, float* colsum, int M, int N, int K,
                     long long gsn, long long gsk, int dtA, vae_stream_t stream);

/* out[M][K] (dtype dtO) = bias[k] + D0.W0 (+ D1.W1), D [M][N] fp32.
 * replaces: aten::addmm at models.py:162 and the input gradient of fc_mu + fc_var */
int vae_skinny_expand(const float* D0, const float* W0, const float* D1, const float* W1,
                      const float* bias, void* out, int M, int N, int K, long long wsn,
                      long long wsk, int dtO, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Layout glue between the NHWC conv stack and the reference's C-major
 * flatten / view (models.py:133 and :166).
 * ---------------------------------------------------------------------- */
int vae_nhwc_to_nchw(const void* src, void* dst, int B, int HW, int C, int dt_src, int dt_dst,
                     vae_stream_t stream);
int vae_nchw_to_nhwc(const void* src, void* dst, int B, int HW, int C, int dt_src, int dt_dst,
                     vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Linear layers fc_mu / fc_var / decoder_input (models.py:55-59,137,141,162)
 * and their gradients: one strided GEMM  C[M,N] (+)= A[M,K] * B[K,N] (+ bias[N]).
 * Element (i,j) of an operand lives at base + i*s_row + j*s_col (in elements).
 * replaces: aten::addmm and its two backward matmuls
 * ---------------------------------------------------------------------- */
int vae_gemm(const void* A, const void* Bm, void* C, const float* bias, int M, int N, int K,
             long long sam, long long sak, long long sbk, long long sbn, long long scm,
             long long scn, int dtA, int dtB, int dtC, int accumulate, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Tensor-core Linear kernels for a WIDE latent (L in {64,128,192,256}, batch M <= 64, F >= 4096 a multiple of 64;
 * the long / wide configuration has F = 262 144, L = 256): every call streams the fp32 master weight (or writes
 * the fp32 gradient) once, rounds operands to bf16 while the mma.sync fragments are formed, accumulates in fp32.
 * Layout codes: 0 = [L][F] (fc_mu / fc_var, models.py:55-56), 1 = [F][L] (decoder_input, models.py:59).
 * ---------------------------------------------------------------------- */
int vae_linear_tc_supported(int M, int L, int F);
/* out[M][L] (fp32, overwritten) = A[M][F] (bf16) . W + bias.
 * replaces: aten::addmm of fc_mu / fc_var (models.py:137,141; layout 0) and the latent gradient of
 * decoder_input (train.py:650; layout 1) */
int vae_linear_reduce_tc(const void* A, const float* W, const float* bias, float* out, int M, int L, int F,
                         int w_layout, vae_stream_t stream);
/* out[M][F] (dt_out) = bias[f] + D0[M][L] . W0 (+ D1 . W1), D fp32.
 * replaces: aten::addmm of decoder_input (models.py:162; layout 1) and the input gradient of fc_mu + fc_var
 * (train.py:650; layout 0, both heads in one pass) */
int vae_linear_expand_tc(const float* D0, const float* W0, const float* D1, const float* W1, const float* bias,
                         void* out, int M, int L, int F, int w_layout, int dt_out, vae_stream_t stream);
/* G (fp32, overwritten; layout 0: [L][F], 1: [F][L]) = D[M][L]^T . A[M][F], D fp32, A bf16.
 * replaces: the weight gradients of the three Linear layers (train.py:650) */
int vae_linear_wgrad_tc(const float* D, const void* A, float* G, int M, int L, int F, int g_layout,
                        vae_stream_t stream);

/* out[n] = sum_m A[m,n]  (bias gradients).  replaces: aten::sum in addmm backward */
int vae_colsum(const void* A, float* out, int M, int N, int dtype, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Reparameterisation (models.py:177-183) with Philox4x32-10 RNG.
 * ---------------------------------------------------------------------- */
/* eps ~ N(0,1), reproducible in (seed, offset); element i uses counter offset+i/4. */
int vae_philox_randn(float* out, long long n, unsigned long long seed,
                     unsigned long long offset, vae_stream_t stream);
/* z = mu + eps*exp(0.5*logvar).  eps_in != NULL injects eps (parity tests);
 * otherwise eps is drawn from Philox and written to eps_out for the backward.
 * dev_offset (nullable) is a device counter added to `offset`, so a captured
 * CUDA graph draws fresh noise on every replay. */
int vae_reparam_fwd(const float* mu, const float* logvar, const float* eps_in, float* eps_out,
                    float* z, long long n, unsigned long long seed, unsigned long long offset,
                    const long long* dev_offset, vae_stream_t stream);
/* dmu = dz ; dlogvar = dz*eps*0.5*exp(0.5*logvar)   (accumulate flag adds). */
int vae_reparam_bwd(const float* dz, const float* eps, const float* logvar, float* dmu,
                    float* dlogvar, long long n, int accumulate, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Fused ELBO (models.py:190-225): mean BCE on sigmoid(logits) with the
 * reference's log clamp at -100, KL term, and all three LossOutput scalars.
 * losses[0]=loss, [1]=reconstruction_loss, [2]=kld_loss (sign-flipped, :224).
 * `scratch` = 4 doubles of workspace (zeroed by this call).
 * input_is_prob != 0: `logits` holds probabilities (a hand-built ModelOutput).
 * replaces: aten::sigmoid (:81), aten::binary_cross_entropy (:208), the
 * sum/mean/exp/pow chain (:214) and the add (:216)
 * ---------------------------------------------------------------------- */
int vae_elbo_fwd(const void* logits, const float* target, const float* mu, const float* logvar,
                 float* losses, double* scratch, long long n_elem, int B, int L,
                 float kld_weight, int dtype, int input_is_prob, vae_stream_t stream);
/* Gradient of losses[0] scaled by *gscale (device scalar, the upstream grad):
 * dlogits = g*(p-x)/n_elem (clamp/saturation emulated), dmu = g*kw*mu/B,
 * dlogvar = g*kw*0.5*(exp(lv)-1)/B. */
int vae_elbo_bwd(const void* logits, const float* target, const float* mu, const float* logvar,
                 const float* gscale, void* dlogits, float* dmu, float* dlogvar,
                 long long n_elem, int B, int L, float kld_weight, int dtype, int input_is_prob,
                 vae_stream_t stream);

/* Forward and gradient of the ELBO in ONE pass over logits + target (north-star item 2): writes
 * the three losses like vae_elbo_fwd and dlogits / dmu / dlogvar like vae_elbo_bwd for an upstream
 * gradient of 1.  bf16 logits use MUFU exp / log / reciprocal (same clamps and saturation points);
 * fp32 logits use the exact functions of the two calls above.
 * replaces: models.py:208-216 plus their autograd backward nodes (train.py:650) */
int vae_elbo_fused(const void* logits, const float* target, const float* mu, const float* logvar,
                   float* losses, double* scratch, void* dlogits, float* dmu, float* dlogvar,
                   long long n_elem, int B, int L, float kld_weight, int dtype, int input_is_prob,
                   vae_stream_t stream);
/* dlogits, dmu, dlogvar *= *gscale in place (the upstream gradient of train.py:650's backward());
 * a no-op launch when *gscale == 1. */
int vae_elbo_scale_grads(const float* gscale, void* dlogits, float* dmu, float* dlogvar,
                         long long n_elem, long long n_lat, int dtype, vae_stream_t stream);

/* probs(fp32) = sigmoid(logits).  replaces: aten::sigmoid at models.py:81 */
int vae_sigmoid(const void* logits, float* probs, long long n, int dtype, vae_stream_t stream);
/* dlogits (+)= dprobs * p * (1-p)  (for user losses defined on probabilities). */
int vae_sigmoid_bwd(const float* dprobs, const float* probs, void* dlogits, long long n,
                    int dtype, int accumulate, vae_stream_t stream);

/* ------------------------------------------------------------------------
 * Device-side piano-roll transforms (data_transformations.py:48-72 "midi"
 * stack: crop -> ToDtype(scale 1/255) -> Normalize(mean,std); plus the
 * binarise / pitch-transpose augmentation north_star item (3) adds).
 * rolls: uint8 [B, Hs, Ws]; out: fp32 [B,1,H,W].
 * crop_mode 0 = center (test transform, :64), 1 = random (train, :54) drawn
 * from Philox(seed, offset); max_pitch_shift > 0 rolls pitches by a random
 * amount in [-max, max] (zero fill); binarize_threshold >= 0 maps the scaled
 * value to {0,1} before normalisation.
 * ---------------------------------------------------------------------- */
int vae_roll_transform(const uint8_t* rolls, float* out, int B, int Hs, int Ws, int H, int W,
                       int crop_mode, int max_pitch_shift, float binarize_threshold, float mean,
                       float std, unsigned long long seed, unsigned long long offset,
                       vae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VAE_B200_H_ */
