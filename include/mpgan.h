/*
 * mpgan.h -- C ABI of libmpgan_sm100.so: the B200 (sm_100a) kernels behind the GAN hot path of
 * mbrzus/Cross-Modality-Minipig-Gan.
 *
 * The reference has no FFI seam for this path: its arithmetic is PyTorch ops called from plain Python
 * classes.  Each entry point below replaces the torch op(s) named in its comment
 * (file:line into /root/reference).  Conventions:
 *   - extern "C", plain pointers and sizes only; the caller owns every buffer (device memory unless noted).
 *   - activations are channels-last: (N, [D,] H, W, C) with a pixel stride `ld*` in elements (>= C), so a
 *     tensor may be a channel slice of a wider concat buffer.
 *   - conv weights are "OTI": [Cy][tap][Cx] (taps in (kd,kh,kw) order) for the underlying convolution
 *     X-grid -> Y-grid; a ConvTranspose is the same weight used in the Y -> X direction.
 *   - dtype: MPGAN_F32 or MPGAN_BF16 for activations/weights; statistics, losses, master weights and
 *     gradients of parameters are always fp32 (batch-norm sums fp64).
 *   - every call is asynchronous on `stream` (a cudaStream_t), never synchronises or allocates, and is
 *     CUDA-graph-capture safe.  Return 0 on success; <0 on error, text via mpgan_last_error().
 *   - no CPU fallback exists: on a machine without a B200 the compute entry points fail with MPGAN_ERR_CUDA.
 */
#ifndef MPGAN_H_
#define MPGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPGAN_VERSION 100

enum { MPGAN_OK = 0, MPGAN_ERR_SHAPE = -1, MPGAN_ERR_UNSUPPORTED = -2, MPGAN_ERR_CUDA = -3 };
enum { MPGAN_F32 = 0, MPGAN_BF16 = 1 };
enum { MPGAN_ACT_NONE = 0, MPGAN_ACT_PRELU = 1, MPGAN_ACT_LEAKY = 2, MPGAN_ACT_TANH = 3 };

/* Geometry of the underlying convolution X-grid -> Y-grid (y = x*stride - pad + r). rank 2 uses index 1,2 of
 * the 3-vectors (index 0: size 1, k 1, stride 1, pad 0). */
typedef struct {
  int32_t rank;       /* 2 or 3 spatial dims */
  int32_t n;          /* batch */
  int32_t xs[3];      /* X-grid spatial extent (D,H,W) */
  int32_t ys[3];      /* Y-grid spatial extent */
  int32_t cx, cy;     /* channels on the X / Y side */
  int32_t k[3], stride[3], pad[3];
} MpganConvGeom;

int mpgan_version(void);
const char* mpgan_last_error(void);
/* 1 if a CUDA device of compute capability 10.x is usable by this process, else 0 (never raises). */
int mpgan_device_ok(void);

/* ---- convolution, generic CUDA-core path (any channel count, rank 2/3, f32 or bf16 storage, fp32 accumulate) ----
 * nn.Conv3d/Conv2d forward  (GAN_final.py:167-189, MONAI Convolution via GAN_final.py:106-114)  = fprop
 * nn.ConvTranspose forward  (MONAI up path)                                                       = bprop
 * and their data gradients (torch autograd's convolution_backward) the other way round.
 * `w` is OTI [cy][taps][cx] in `dtype`; bias fp32 [channels of the output side] or NULL. */
int mpgan_conv_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w,
                     const float* bias, void* y, int64_t ldy, void* stream);
int mpgan_conv_bprop(const MpganConvGeom* g, int dtype, const void* y, int64_t ldy, const void* w,
                     const float* bias, void* x, int64_t ldx, void* stream);
/* dw[cy][taps][cx] += sum_pixels y (x) x   (fp32, accumulates: the caller zeroes);  weight gradient of both
 * Conv and ConvTranspose. */
int mpgan_conv_wgrad(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* y, int64_t ldy,
                     float* dw, void* stream);

/* ---- convolution, tcgen05 path (bf16 operands, fp32 TMEM accumulators, TMA-staged NHWC tiles; rank 2) ----
 * Same contracts as above with dtype == BF16.  Requirements: cx, cy multiples of 16 (>= 16).
 * w_f is OTI [cy][taps][cx] bf16 (used by fprop); w_b is its transpose [cx][taps][cy] bf16 (used by bprop).
 * Optional fused epilogue: bias add and per-channel batch-norm partial sums (sum, sum of squares of the
 * stored bf16 values) accumulated into stats[2*C] (fp64, caller zeroes) when stats != NULL. */
int mpgan_tc_supported(const MpganConvGeom* g, int direction /*0 fprop,1 bprop,2 wgrad*/);
int mpgan_tc_conv_fprop(const MpganConvGeom* g, const void* x, int64_t ldx, const void* w_f, const float* bias,
                        void* y, int64_t ldy, double* stats, void* stream);
int mpgan_tc_conv_bprop(const MpganConvGeom* g, const void* y, int64_t ldy, const void* w_b, const float* bias,
                        void* x, int64_t ldx, double* stats, void* stream);
/* same, with a bf16 tensor `res` (pixel stride ldres, the X grid) added to the result before rounding: fuses the
 * gradient accumulation of residual / skip branches (dx = dgrad(dy) + res) into the convolution's epilogue */
/* inference-mode fused layer (BatchNorm folded into w / bias by the caller):  out = prelu(conv(in, w) + bias) + res
 * direction 0: convolution X -> Y with w = [cy][taps][cx]; direction 1: transposed convolution Y -> X with the
 * transposed weights [cx][taps][cy].  slope: device scalar (PReLU weight) or NULL (no activation); res optional. */
int mpgan_tc_conv_act(const MpganConvGeom* g, int direction, const void* in, int64_t ldi, const void* w,
                      const float* bias, const float* slope, const void* res, int64_t ldres, void* out, int64_t ldo,
                      void* stream);
/* ConvTranspose(cy -> 1, k3 s2 p1 op1) forward == data gradient of a one-input-channel stride-2 3x3 convolution, through
 * the halo tcgen05 kernel in pixel-shuffle mode: w = the layer's bf16 [cy][9] weight, x = (n, 2*yh, 2*yw) one channel,
 * stats (optional) = fp64 {sum, sum of squares} of the stored values (fused BatchNorm(1) statistics). */
int mpgan_tc_convt_to1(const MpganConvGeom* g, const void* y, int64_t ldy, const void* w, const float* bias, void* x,
                       double* stats, void* stream);
/* data gradient of a one-input-channel stride-1 3x3 layer (cx == 1; dY has cy in {16,32,64,128} channels) through the
 * halo-resident tcgen05 kernel: w_b16 = bf16 [16][9][cy], row 0 the layer's transposed weights, rows 1..15 zero;
 * x: (n, xh, xw) one channel, pixel stride ldx; res (optional, same layout, stride ldres) is added. */
int mpgan_tc_conv_bprop_c1out(const MpganConvGeom* g, const void* y, int64_t ldy, const void* w_b16, void* x,
                              int64_t ldx, const void* res, int64_t ldres, void* stream);
int mpgan_tc_conv_bprop_res(const MpganConvGeom* g, const void* y, int64_t ldy, const void* w_b, const float* bias,
                            void* x, int64_t ldx, const void* res, int64_t ldres, double* stats, void* stream);
/* workspace_bytes from mpgan_tc_conv_wgrad_workspace(); dw accumulates (fp32 [cy][taps][cx]). */
size_t mpgan_tc_conv_wgrad_workspace(const MpganConvGeom* g);
int mpgan_tc_conv_wgrad(const MpganConvGeom* g, const void* x, int64_t ldx, const void* y, int64_t ldy, float* dw,
                        void* workspace, size_t workspace_bytes, void* stream);
/* ---- convolution, one-channel edge layers: bandwidth-bound direct kernels, rank 2/3, f32 or bf16 ----
 * fprop with cy == 1 or (rank 2, 3x3) cx == 1 (G's 1->16 and D's 1->64 first convolutions), bprop with cx == 1 (ConvTranspose 32->1
 * forward, data gradient of the 1->16 / 1->64 first convolutions), wgrad with cx == 1.  Same contracts as the generic
 * entry points; mpgan_c1_supported() says whether a (geometry, direction) is covered. */
int mpgan_c1_supported(const MpganConvGeom* g, int direction /*0 fprop,1 bprop,2 wgrad*/);
int mpgan_c1_conv_fprop(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w,
                        const float* bias, void* y, int64_t ldy, double* stats /* nullable: BN sums of y */,
                        void* stream);
/* inference fusion of a one-input-channel layer: y = prelu(conv(x, w) + bias), BatchNorm folded into (w, bias) by the caller
 * (MONAI Convolution in model.eval(), inferrence.py:107-109); returns -2 when the run-based kernel does not cover the layer */
int mpgan_c1_conv_act(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* w, const float* bias,
                      const float* slope, void* y, int64_t ldy, void* stream);
int mpgan_c1_conv_bprop(const MpganConvGeom* g, int dtype, const void* y, int64_t ldy, const void* w,
                        const float* bias, void* x, int64_t ldx, double* stats /* nullable: BN sums of x */,
                        void* stream);
int mpgan_c1_conv_wgrad(const MpganConvGeom* g, int dtype, const void* x, int64_t ldx, const void* y, int64_t ldy,
                        float* dw, void* stream);

/* ---- batch norm (nn.BatchNorm3d, GAN_final.py:170-188; MONAI norm=BATCH) + activation, fused ----
 * stats: per-channel sum / sum-of-squares in fp64 (accumulates; caller zeroes). */
int mpgan_bn_stats(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, double* stats, void* stream);
/* training: batch mean / biased var -> scale=gamma*invstd, shift=beta-mean*scale; saves mean, invstd;
 * running_mean/var updated with `momentum` (unbiased var), num_batches_tracked (int64) += 1.
 * eval (training==0): scale/shift from the running statistics; stats may be NULL. */
int mpgan_bn_finalize(const double* stats, int64_t pixels, int32_t c, const float* gamma, const float* beta,
                      float eps, float momentum, int training, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                      void* stream);
/* y = act(x*scale[c] + shift[c]) (+ res).  act: NONE, PRELU (slope read from *alpha), LEAKY (slope = *alpha
 * when alpha != NULL else 0.2), TANH.  scale/shift may be NULL (identity).  res may be NULL. */
int mpgan_bn_act_apply(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, const float* scale,
                       const float* shift, int act, const float* alpha, float leaky_slope, const void* res,
                       int64_t ldres, void* y, int64_t ldy, void* stream);
/* training-mode finalize + apply in one launch: every thread derives scale/shift of its own channels from the fp64
 * statistics (the arithmetic of mpgan_bn_finalize); the first block also writes mean/invstd/scale/shift (saved for
 * the backward pass) and updates running_mean/var and num_batches_tracked. */
int mpgan_bn_train_apply(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, const double* stats,
                         const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, int64_t* num_batches_tracked, float* mean, float* invstd, float* scale,
                         float* shift, int act, const float* alpha, float leaky_slope, const void* res,
                         int64_t ldres, void* y, int64_t ldy, void* stream);
/* backward of y = act(bn(x)): pass 1 reduces  sums[0:C]=sum g, sums[C:2C]=sum g*xhat, sums[2C]=sum dy*min(z,0)
 * (PReLU slope grad), all fp64 accumulating;  pass 2 writes dx and accumulates dgamma/dbeta/dalpha (fp32) and, when
 * dbias != NULL, the bias gradient of the convolution that produced x: dbias[c] += sum_p dx[p,c]. */
int mpgan_bn_act_bwd_reduce(int dtype, const void* dy, int64_t lddy, const void* x, int64_t ldx, int64_t pixels,
                            int32_t c, const float* mean, const float* invstd, const float* scale,
                            const float* shift, int act, const float* alpha, float leaky_slope, double* sums,
                            void* stream);
int mpgan_bn_act_bwd_apply(int dtype, const void* dy, int64_t lddy, const void* x, int64_t ldx, int64_t pixels,
                           int32_t c, const float* mean, const float* invstd, const float* scale,
                           const float* shift, int act, const float* alpha, float leaky_slope, const double* sums,
                           float* dgamma, float* dbeta, float* dalpha, float* dbias, void* dx, int64_t lddx,
                           void* stream);

/* ---- elementwise helpers on channels-last tensors ---- */
/* y[p, 0:c] = a[p, 0:c] (+ b[p, 0:c]) with independent pixel strides (residual add, concat copy, casts). */
int mpgan_add_copy(int dtype_in, const void* a, int64_t lda, const void* b, int64_t ldb, int dtype_out, void* y,
                   int64_t ldy, int64_t pixels, int32_t c, void* stream);
/* nn.Tanh (GAN_final.py:117) forward and backward on a flat tensor: y = tanh(x);  dx = dy*(1-y^2). */
int mpgan_tanh_fwd(int dtype_in, const void* x, int dtype_out, void* y, int64_t n, void* stream);
int mpgan_tanh_bwd(int dtype, const void* dy, const void* y, void* dx, int64_t n, void* stream);
/* per-channel column sum: out[c] += sum_p x[p,c]  (conv bias gradient). fp32 accumulate into out. */
int mpgan_colsum(int dtype, const void* x, int64_t ldx, int64_t pixels, int32_t c, float* out, void* stream);

/* ---- nn.Linear after Flatten (GAN_final.py:200-201; test_runs/GAN.py:176-181) ----
 * x: (batch, k) in `dtype` (channels-last flatten order), w: (j, k) in `dtype` pre-permuted to the same order,
 * bias fp32 (j) or NULL, y fp32 (batch, j).  fwd accumulates into y (caller zeroes y). */
int mpgan_linear_fwd(int dtype, const void* x, const void* w, const float* bias, float* y, int32_t batch,
                     int64_t k, int32_t j, void* stream);
/* dx (batch,k) `dtype` = dy (batch,j) fp32 * w;   dw (j,k) fp32 += dy^T x;  db (j) += sum dy.  NULL skips. */
int mpgan_linear_bwd(int dtype, const void* x, const void* w, const float* dy, void* dx, float* dw, float* db,
                     int32_t batch, int64_t k, int32_t j, void* stream);
/* (c, spatial) <-> (spatial, c) permutation of each row of a (rows, c*spatial) matrix, with dtype conversion:
 * to_cl=1: dst[r][s][c] = src[r][c][s];  to_cl=0: dst[r][c][s] (+)= src[r][s][c]. */
int mpgan_permute_flatten(int dtype_src, const void* src, int dtype_dst, void* dst, int32_t rows, int32_t c,
                          int64_t spatial, int to_cl, int accumulate, void* stream);

/* ---- losses ----
 * nn.Sigmoid (GAN_final.py:203) on the (n) fp32 logits and its backward dz = dprob * p * (1-p). */
int mpgan_sigmoid_fwd(const float* z, float* prob, int32_t n, void* stream);
int mpgan_sigmoid_bwd(const float* dprob, const float* prob, float* dz, int32_t n, void* stream);
/* F.binary_cross_entropy (GAN_final.py:244-245) on probabilities, torch's clamp semantics:
 * loss += weight * mean(-(t*max(log p,-100) + (1-t)*max(log1p(-p),-100)));  loss (1) fp32 accumulates.
 * backward: dprob = gscale * weight/n * (p - t)/max(p(1-p), 1e-12)   (gscale: device scalar or NULL = 1). */
int mpgan_bce_fwd(const float* prob, const float* target, float weight, float* loss, int32_t n, void* stream);
int mpgan_bce_bwd(const float* prob, const float* target, float weight, const float* gscale, float* dprob,
                  int32_t n, void* stream);
/* F.l1_loss (GAN_final.py:247-248): loss += weight * mean|a-b| ;  da = gscale*weight/n * sign(a-b) */
int mpgan_l1_fwd(int dtype, const void* a, const void* b, int64_t n, float weight, float* loss, void* stream);
int mpgan_l1_bwd(int dtype, const void* a, const void* b, int64_t n, float weight, const float* gscale, void* da,
                 int accumulate, void* stream);
/* one pass: *loss (fp32) and / or *loss64 (fp64) += weight * mean|a - b| (either may be NULL) and
 * da (+)= gscale * weight / n * sign(a - b) (da may be NULL; accumulate != 0 adds onto da's contents) -- F.l1_loss forward +
 * backward fused; the feature-matching loss of test_runs/GAN.py:288-298 accumulates its gradients straight into the
 * discriminator's backward tensors with it and sums its 16 terms in the fp64 slot. */
int mpgan_l1_fwd_bwd(int dtype, const void* a, const void* b, int64_t n, float weight, const float* gscale, float* loss,
                     double* loss64, void* da, int accumulate, void* stream);

/* ---- torch.optim.Adam (GAN_final.py:298-308), one launch over a flat fp32 buffer ----
 * state: 3 device floats {step count (int bits), step_size, sqrt(1-beta2^t)}; zero-initialised by the caller.
 * The call increments the device-side step count first (so a captured CUDA graph advances it on replay), then
 * applies exp_avg/exp_avg_sq/param updates with torch's formula.  Optionally writes a bf16 shadow copy. */
int mpgan_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float* state, void* bf16_shadow, void* stream);
/* dst = cast(src) for flat buffers (f32 <-> bf16). */
int mpgan_cast(int dtype_src, const void* src, int dtype_dst, void* dst, int64_t n, void* stream);
/* OTI [cy][taps][cx] -> transposed shadow [cx][taps][cy] with dtype conversion. */
int mpgan_weight_transpose(int dtype_src, const void* src, int dtype_dst, void* dst, int32_t cy, int32_t taps,
                           int32_t cx, void* stream);
/* the same for n bf16 tensors in one launch: table = n x 5 device int64 {src, dst, cy, taps, cx}; max_elems = the largest
 * cy*taps*cx (grid sizing) */
int mpgan_weight_transpose_batch(const int64_t* table_dev, int32_t n, int32_t max_elems, void* stream);

/* ---- RandSpatialCropSamplesd gather (test_runs/GAN.py:263-272,313-337), bit-exact copy ----
 * vol: (batch, *S, c) channels-last; origins: device int32 (batch*num_samples, rank) in (D,H,W) order;
 * out: (batch*num_samples, roi.., c).  scatter_add is its deterministic backward (dvol += patches). */
int mpgan_patch_gather(int dtype, const void* vol, int32_t batch, int32_t rank, const int32_t* spatial, int32_t c,
                       const int32_t* origins, int32_t num_samples, int32_t roi, void* out, void* stream);
int mpgan_patch_scatter_add(int dtype, const void* dpatch, int32_t batch, int32_t rank, const int32_t* spatial,
                            int32_t c, const int32_t* origins, int32_t num_samples, int32_t roi, void* dvol,
                            void* stream);

/* ---- fused tail of a generator UNet (MONAI UNet top level: ConvT(..->1) -> BatchNorm(1) -> PReLU -> ResidualUnit(1->1,
 * conv only), call site GAN_final.py:106-114) on a one-channel (n, h, w) bf16 image, training mode:
 *   hmap = prelu(batchnorm(c));  y = conv3x3(hmap, w9, pad 1) + bias + hmap
 * stats: the fp64 {sum, sum of squares} of c; running statistics / saved mean, invstd, scale, shift are updated as by
 * mpgan_bn_train_apply.  h_out may be NULL (no backward pass will follow).  stats == NULL: evaluation mode -- scale /
 * shift are INPUTS (mpgan_bn_finalize of the running statistics) and no statistic is written. */
int mpgan_c1_tail_fwd(const void* c_bf16, int32_t n, int32_t h, int32_t w, const double* stats, const float* gamma,
                      const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, float* mean, float* invstd, float* scale, float* shift,
                      const float* alpha, const void* w9_bf16, const float* bias, void* h_out_bf16, void* y_out_bf16,
                      void* stream);
/* backward twin: dh = conv3x3^T(dy, w9) + dy (bf16, same layout) and the BatchNorm(1) + PReLU backward reduction of c in
 * the same pass: sums3[0] += sum dz, sums3[1] += sum dz*xhat, sums3[2] += sum_{z<=0} dh*z (the layout
 * mpgan_bn_act_bwd_apply reads with c == 1); mean / invstd / scale / shift = the values saved by the forward. */
int mpgan_c1_tail_bwd_reduce(const void* dy_bf16, const void* c_bf16, int32_t n, int32_t h, int32_t w, const float* mean,
                             const float* invstd, const float* scale, const float* shift, const float* alpha,
                             const void* w9_bf16, void* dh_bf16, double* sums3, void* stream);

/* ---- one-input-channel 3x3 weight gradient through the tensor cores (conv_c1col.cu) ----
 * im2col_c1: x (n, ih, iw) contiguous bf16, one channel -> xcol (n, oh, ow, 16) bf16: the 9 taps of every output pixel
 *   (zero padded borders) followed by 7 zeros.  The weight gradient is then mpgan_tc_conv_wgrad of the 1x1 layer
 *   xcol(16) -> dY(cy) into a zeroed fp32 dw16[cy][16], and fold_dw16 adds dw16[c][t<9] into dw[c][9]. */
int mpgan_im2col_c1(const void* x_bf16, int32_t n, int32_t ih, int32_t iw, int32_t oh, int32_t ow, int32_t stride,
                    int32_t pad, void* xcol_bf16, void* stream);
int mpgan_fold_dw16(const float* dw16, int32_t cy, float* dw, void* stream);

/* ---- one-channel layers of the rank-3 networks (conv_c1vol.cu; the reference's literal volumes: GAN_final.py:107
 * dimensions=3, :167-169 Conv3d(1, 64, 3), MONAI UNet's 1 -> 16 entry convolutions / 32 -> 1 ConvTranspose / 1 -> 1 tail) ----
 * A 3x3x3 convolution with one input channel == a 1x1x1 convolution over xcol (n, *ys, 32) = the 27 taps of every output
 * voxel + 5 zeros, which the rank-3 tcgen05 kernels run (mpgan_tc_conv_fprop / _wgrad with cx = 32):
 * im2col_c1_vol: x (n, *xs3) one channel contiguous bf16 -> xcol (n, *ys3, 32) bf16 (zero padded borders).
 * col2im_c1_vol: t (n, *ys3, 32) bf16 = per-tap partial products of a data gradient / ConvTranspose (the 1x1x1 convolution
 *   of the Y-grid tensor with the [32][cy] transposed weights) -> x (n, *xs3) one channel bf16:
 *   x[q] = bias + res[q] + sum over taps r with (q + pad - r) % stride == 0 of t[(q + pad - r) / stride][r].
 * fold_dw32: dw[c][27] += dw32[c][32] (first 27 columns of the 1x1x1 weight gradient).
 * stencil27 / stencil27_wgrad: the 1 -> 1 channel k3 s1 p1 layer as direct 27-point stencils (dtype MPGAN_F32 / MPGAN_BF16;
 *   direction 0: y = bias + conv(x, w27); 1: dx = conv^T(dy, w27) + res; wgrad: dw27[t] += sum dy * x(shifted), dbias += sum dy). */
int mpgan_im2col_c1_vol(const void* x_bf16, int32_t n, const int32_t* xs3, const int32_t* ys3, int32_t stride, int32_t pad,
                        void* xcol_bf16, void* stream);
int mpgan_col2im_c1_vol(const void* t_bf16, int32_t n, const int32_t* xs3, const int32_t* ys3, int32_t stride, int32_t pad,
                        const float* bias, const void* res_bf16, void* x_bf16, void* stream);
int mpgan_fold_dw32(const float* dw32, int32_t cy, float* dw, void* stream);
int mpgan_stencil27(int dtype, int direction, const void* in, int32_t n, int32_t d, int32_t h, int32_t w, const void* w27,
                    const float* bias, const void* res, void* out, void* stream);
int mpgan_stencil27_wgrad(int dtype, const void* x, const void* dy, int32_t n, int32_t d, int32_t h, int32_t w, float* dw27,
                          float* dbias, void* stream);

/* ---- intensity transforms / volume metrics around the generator ----
 * MONAI ScaleIntensityRangePercentilesd (reference: GAN_final.py:386-394 lower=1 upper=99 -> [-1,1];
 * inferrence.py:152-160,190-198 lower=0 upper=100 -> [0,255] followed by np.round) and torchmetrics
 * MeanAbsoluteError / MeanSquaredError (inferrence.py:170-176, metrics.py:213-218).
 * order_stats: out[r] = the ranks[r]-th smallest value (0-based, exact) of the fp32 volume x[0..n); ranks is a DEVICE
 *   array of nranks <= 4 int64; workspace of mpgan_order_stats_workspace(nranks) bytes (caller-owned, overwritten).
 * minmax: the same contract restricted to ranks in {0, n-1} (the 0 / 100 percentiles): one min/max reduction pass.
 * rescale_intensity: y = ((x - a_min) / (a_max - a_min)) * (b_max - b_min) + b_min  (x - a_min + b_min when
 *   a_max == a_min), each step rounded to fp32; optional clip to [clip_lo, clip_hi]; optional round-half-even;
 *   out_dtype 0 = fp32, 2 = fp16.
 * err_sums: out2[0] += sum |a-b|, out2[1] += sum (a-b)^2  (device fp64, zero-initialised by the caller). */
size_t mpgan_order_stats_workspace(int32_t nranks);
int mpgan_order_stats(const float* x, int64_t n, const int64_t* ranks, int32_t nranks, float* out, void* workspace,
                      size_t workspace_bytes, void* stream);
int mpgan_minmax(const float* x, int64_t n, const int64_t* ranks, int32_t nranks, float* out, void* workspace,
                 size_t workspace_bytes, void* stream);
int mpgan_rescale_intensity(const float* x, int64_t n, float a_min, float a_max, float b_min, float b_max, int clip,
                            float clip_lo, float clip_hi, int round_half_even, int out_dtype, void* y, void* stream);
int mpgan_err_sums(const float* a, const float* b, int64_t n, double* out2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPGAN_H_ */
