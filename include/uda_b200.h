/* uda_b200 — C ABI of the B200-native (sm_100a) segmentation-training hot path.
 *
 * Drop-in boundary for bempt/uda_aerial_semantic_segmentation_research (a pure-Python PyTorch repo
 * with NO native/FFI layer of its own: SURVEY.md 8b).  Each entry point below replaces the ATen /
 * cuDNN library calls the reference reaches through the cited Python call site; the Python shims in
 * uda_aerial_semantic_segmentation_research_b200/ bind them with ctypes (INTEGRATION.md shows the
 * stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every buffer is BORROWED device memory owned by the caller
 *     (torch caching allocator); kernels never allocate or free; scratch is an explicit
 *     `workspace` argument whose size is documented per call;
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); no internal
 *     synchronisation; safe under CUDA-graph capture;
 *   - return 0 on success or a negative UDA_ERR_* code; uda_last_error() returns the
 *     thread-local message; arguments are validated BEFORE any launch; there is no CPU fallback;
 *   - dtype codes: UDA_F32 / UDA_BF16 for activations and logits, UDA_I64 / UDA_U8 for index maps;
 *   - activations inside the network are NHWC ([B,H,W,C], C innermost); logits / losses use the
 *     reference's NCHW ([B,C,H,W]) edge layout; conv weights are OHWI ([Cout][KH][KW][Cin]),
 *     the physical layout of a torch channels_last parameter of logical shape [Cout,Cin,KH,KW].
 */
#ifndef UDA_B200_H_
#define UDA_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UDA_OK 0
#define UDA_ERR_BAD_ARG (-1)
#define UDA_ERR_UNSUPPORTED (-2)
#define UDA_ERR_CUDA (-3)

#define UDA_F32 0
#define UDA_BF16 1
#define UDA_I64 2
#define UDA_U8 3

/* library */
const char* uda_last_error(void);
int uda_abi_version(void);
/* 1 if the current device is sm_100 (B200) and the tcgen05 kernels can run, else 0 */
int uda_device_supported(void);

/* ---------------------------------------------------------------------------------------------
 * Loss path (fused loss + logit gradient).
 * ------------------------------------------------------------------------------------------- */

/* Segmentation loss:  total = out_scale * (w_ce * CE_term + w_dice * Dice_term), gradient wrt logits.
 *   ce_mode 0: no CE term; 1: nn.CrossEntropyLoss (reference src/models/train.py:208,342;
 *   adversarial_trainer.py:105) mean over non-ignored pixels, optional class weights;
 *   2: focal(weighted CE) of WeightedSegmentationLoss.focal_loss (src/models/losses.py:176-187).
 *   use_dice: DiceLoss.forward (src/models/losses.py:118-152) on index targets, or on
 *   `soft_target` ([B,C,HW] float one-hot/soft, as losses.py:134 accepts) when non-NULL.
 *   logits/grad: [B,C,HW] contiguous, `dtype`; target: int64 [B,HW]; class_weights: float[C] or NULL.
 *   out4 (device float[4]) = {CE term, Dice term, total, #out-of-range targets}.
 *   Single pass when use_dice == 0, two passes (sums, then gradient) otherwise.
 *   workspace: uda_seg_loss_workspace_bytes(B, C), 8-byte aligned. */
size_t uda_seg_loss_workspace_bytes(int B, int C);
int uda_seg_loss_fwd_bwd(const void* logits, int dtype, const long long* target, const float* soft_target,
                         const float* class_weights, void* grad, float* out4, void* workspace, int B, int C,
                         long long HW, int ce_mode, int use_dice, float alpha, float gamma, int mean_reduction,
                         long long ignore_index, float smooth, float w_ce, float w_dice, float out_scale,
                         void* stream);

/* x[i] *= *dev_scalar, skipped entirely when *dev_scalar == 1 (autograd grad_output hook). */
int uda_scale_by_device_scalar(void* x, int dtype, long long n, const float* dev_scalar, void* stream);

/* ConsistencyLoss.forward (src/models/losses.py:62-90): out1[0] = out_scale * symmetric KL
 * (temperature, 'batchmean'); grad1/grad2 = d/dz1, d/dz2.  workspace: 8 bytes. */
int uda_consistency_fwd_bwd(const void* z1, const void* z2, int dtype, void* grad1, void* grad2, float* out1,
                            void* workspace, int B, int C, long long HW, float temperature, float out_scale,
                            void* stream);

/* Target-domain entropy minimisation (north-star extension):
 * out1[0] = out_scale * mean_px(-sum_c p log p).  workspace: 8 bytes. */
int uda_entropy_fwd_bwd(const void* z, int dtype, void* grad, float* out1, void* workspace, int B, int C,
                        long long HW, float out_scale, void* stream);

/* nn.BCEWithLogitsLoss() against a constant label (AdversarialLoss, src/models/losses.py:16-51):
 * out1[0] (+)= scale * mean BCE(x, label); grad (nullable) = d/dx. */
int uda_bce_logits_fwd_bwd(const float* x, float* grad, float* out1, long long n, float label, float scale,
                           int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Prediction / evaluation (integer results, bit-exact).
 * ------------------------------------------------------------------------------------------- */

/* outputs.argmax(dim=1) (src/models/predict.py:129, src/models/train.py:227) fused with
 * SegmentationMetrics._fast_hist (src/analysis/metrics.py:17-27).  logits [B,C,HW]; target int64
 * [B,HW] or NULL; mask_i64 / mask_u8 / hist ([C,C] int64, rows=true, cols=pred) each nullable. */
int uda_argmax_confmat(const void* logits, int dtype, const long long* target, long long* mask_i64,
                       unsigned char* mask_u8, long long* hist, int B, int C, long long HW,
                       long long ignore_index, int has_ignore, int zero_hist, void* stream);

/* SegmentationMetrics._fast_hist on two index maps (pred int64 or uint8). bad_count (nullable)
 * receives the number of predictions outside [0,C). */
int uda_confmat(const void* pred, int pred_dtype, const long long* target, long long* hist, long long* bad_count,
                long long n, int C, long long ignore_index, int has_ignore, int zero_hist, void* stream);

/* ---------------------------------------------------------------------------------------------
 * U-Net / discriminator layers (replace aten::_convolution, batch_norm, relu, max_pool2d,
 * upsample_nearest2d + cat reached from smp.Unet — src/models/train.py:572-577 — and
 * DomainDiscriminator — src/models/discriminator.py:15-55).
 * ------------------------------------------------------------------------------------------- */
int uda_nchw_f32_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int Cpad, long long HW, void* stream);
int uda_nhwc_to_nchw_f32(const void* src, int dtype, float* dst, int B, int C, int Cpad, long long HW, void* stream);
int uda_cast_f32(const float* src, void* dst, int dtype, long long n, void* stream);

/* Generic FP32-pipe implicit GEMM (any k/stride/pad/channels; fp32 parity mode and odd shapes). */
int uda_conv2d_direct_fwd(const void* x, int dtype, const void* w, int w_dtype, const float* bias, void* y_nhwc,
                          float* y_nchw_f32, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                          int pad, void* stream);
/* dgrad: dx = conv_transpose(dy, w) (+ addend: same shape as dx, may alias dx — gradient accumulation
 * for tensors with several consumers: residual identity, encoder skip connections) */
int uda_conv2d_direct_dgrad(const void* dy, int dtype, const void* w, int w_dtype, const void* addend, void* dx,
                            int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream);
int uda_conv2d_direct_wgrad(const void* dy, const void* x, int dtype, float* dw, int B, int H, int W, int Cin,
                            int Cout, int KH, int KW, int stride, int pad, void* stream);

/* tcgen05 / TMEM / TMA implicit-GEMM convolution (bf16 NHWC in, fp32 accumulate).
 * uda_conv2d_tc_supported returns 1 when the shape is covered (else use the direct kernels).
 *   op 0: fwd   y = conv(x, w)                    x [B,H,W,Cin]  w OHWI bf16  y [B,Ho,Wo,Cout]
 *   op 1: dgrad dx = conv_transpose(dy, w)
 *   op 2: wgrad dw += dy^T * im2col(x)            dw fp32 OHWI
 * Epilogue (fwd): optional bias[Cout]; optional fp32 NCHW second output (logits edge);
 * optional per-channel sum / sum-of-squares accumulation (BatchNorm batch statistics, double[2*Cout]). */
int uda_conv2d_tc_supported(int op, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad);
int uda_conv2d_tc_fwd(const void* x, const void* w, const float* bias, void* y_nhwc, float* y_nchw_f32,
                      double* bn_sums, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                      void* stream);
/* Inference form of the forward convolution: y = act(conv(x, w) + bias (+ addend)) in one launch; act_slope 0 = ReLU,
 * 0.2 = LeakyReLU, 1 = none.  With uda_bn_fold_conv (w' = w * gamma / sqrt(var + eps) in bf16, bias' = beta + (b - mean)
 * * gamma / sqrt(var + eps); w fp32 [Cout][per_out]) an eval-mode conv + BatchNorm (+ residual) + ReLU chain
 * (reference src/models/predict.py:113-130: model.eval(); model(images)) runs without any BatchNorm pass. */
int uda_conv2d_tc_fwd_fused(const void* x, const void* w, const float* bias, const void* addend, float act_slope,
                            void* y_nhwc, float* y_nchw_f32, int B, int H, int W, int Cin, int Cout, int KH, int KW,
                            int stride, int pad, void* stream);
int uda_bn_fold_conv(const float* w, const float* conv_bias, const float* gamma, const float* beta,
                     const float* running_mean, const float* running_var, float eps, void* w_folded, float* bias_folded,
                     int Cout, int per_out, void* stream);
/* Output-space adversarial path (north-star config 3): softmax(logits) packed as the discriminator's channels-last bf16
 * operand (channels >= C zero, Cpad in {8,16,32}); its backward with the gradient-reversal factor folded in
 * (dlogits (+)= scale * p_c * (dp_c - sum_k p_k dp_k), scale = -alpha; reference GradientReverseFunction
 * src/models/uda.py:103-112); y = scale * x (the stand-alone gradient-reversal backward); channel pad / un-pad glue for
 * a first-layer weight whose input channel count is the class count. */
int uda_softmax_nchw_to_nhwc(const float* logits, void* probs, int B, int C, int Cpad, long long HW, void* stream);
int uda_softmax_bwd_grl(const void* probs, const void* dprobs, float* dlogits, float scale, int accumulate, int B, int C,
                        int Cpad, long long HW, void* stream);
int uda_scale(const void* x, void* y, int dtype, float scale, long long n, void* stream);
int uda_pad_channels(const void* src, void* dst, long long rows, int c, int cpad, void* stream);
int uda_unpad_channels_add(const float* src, float* dst, long long rows, int c, int cpad, void* stream);
/* Decoder conv1 of the U-Net WITHOUT the upsampled / concatenated tensor (reference: smp's DecoderBlock
 * F.interpolate(x, 2, 'nearest') -> torch.cat([x, skip], 1) -> conv3x3, created at src/models/train.py:572-577):
 *   conv3x3(cat(up2(x), skip), W) = conv_transpose4x4_s2_p1(x, W4) + conv3x3(skip, Ws)
 * uda_upconv_split_weights: w bf16 [O][3][3][C1+C2] -> wx_ft bf16 [O][4][4][C1] (tap groups summed) and ws bf16 [O][3][3][C2];
 * uda_upconv_tc_fwd: y[B,H,W,O] = act(conv_transpose(x[B,H/2,W/2,C1], wx_ft) + bias (+ addend)) (+ BatchNorm statistics);
 * uda_conv2d_tc_fwd_add: y = conv(x, w) + addend with the BatchNorm statistics of the sum (the skip half);
 * uda_upconv_merge_wgrad: dW fp32 [O][3][3][C1+C2] += un-grouped dW4 fp32 [C1][4][4][O] (x channels), dWs (skip channels).
 * Backward of the x half = the forward / wgrad entry points of the 4x4 stride-2 convolution with W4 =
 * uda_conv2d_weight_flip_transpose(wx_ft). */
int uda_upconv_split_weights(const void* w, void* wx_ft, void* ws, void* w4, void* ws_ft, int Cout, int C1, int C2,
                             void* stream);   /* w4 / ws_ft (nullable): the flipped-transposed copies for the backward */
int uda_upconv_tc_fwd(const void* x, const void* wx_ft, const float* bias, const void* addend, float act_slope, void* y,
                      double* bn_sums, int B, int H, int W, int C1, int Cout, void* stream);
int uda_conv2d_tc_fwd_add(const void* x, const void* w, const void* addend, void* y_nhwc, double* bn_sums, int B, int H,
                          int W, int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream);
int uda_upconv_merge_wgrad(const float* dw4, const float* dws, float* dw, int Cout, int C1, int C2, void* stream);
/* Training-mode conv + BatchNorm2d + activation (+ residual add) as ONE launch (replaces aten::cudnn_convolution ->
 * aten::batch_norm -> aten::add_ -> aten::relu_ of a torchvision BasicBlock / smp Conv2dReLU under the model created at
 * src/models/train.py:572-577): z = conv(x, w) (+ addend), batch statistics of the bf16-rounded z, grid-wide barrier
 * with the accumulators held in TMEM, a = act(z*scale + shift (+ residual)); z and a are both written (z is saved for
 * the backward), mean / rstd / scale / shift and the running statistics as uda_bn_apply_fused.  bn_sums (double[2*Cout])
 * and counter (one uint32) must be zero on entry.  Returns UDA_ERR_UNSUPPORTED before launching anything when the
 * layer's output tiles do not fit the tensor memory of one wave of CTAs (callers fall back to uda_conv2d_tc_fwd +
 * uda_bn_apply_fused).  All CTAs of the launch spin on the barrier: do not run two such launches concurrently on
 * different streams of one device. */
int uda_conv2d_tc_fwd_bn_act(const void* x, const void* w, const void* addend, const void* residual, void* z, void* a,
                             double* bn_sums, unsigned int* counter, const float* gamma, const float* beta,
                             float* running_mean, float* running_var, float* mean, float* rstd, float* scale,
                             float* shift, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
                             float eps, float momentum, float slope, void* stream);
/* dgrad takes w_ft = uda_conv2d_weight_flip_transpose(w): [Cin][KH][KW][Cout] bf16 (the weights of the
 * equivalent forward convolution of dy); addend as in uda_conv2d_direct_dgrad. */
int uda_conv2d_weight_flip_transpose(const void* w, void* w_ft, int Cout, int Cin, int KH, int KW, void* stream);
/* the same for every conv weight of a network in one launch: both bases are bf16 buffers with identical offsets;
 * table_dev = device int[n_weights][5] = {element offset, Cout, Cin, KH, KW} */
int uda_conv2d_weight_flip_transpose_batch(const void* w_base, void* w_ft_base, const int* table_dev, int n_weights,
                                           void* stream);
int uda_conv2d_tc_dgrad(const void* dy, const void* w_ft, const void* addend, void* dx, int B, int H, int W, int Cin,
                        int Cout, int KH, int KW, int stride, int pad, void* stream);
int uda_conv2d_tc_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int KH,
                        int KW, int stride, int pad, void* stream);

/* Cin = 3 stems (U-Net 7x7 s2 p3, discriminator 4x4 s2 p1) on the tensor cores: the fp32 NCHW image is repacked
 * into a zero-padded 4-channel bf16 buffer (uda_stem_packed_input_elems elements) whose kernel rows are K chunks of
 * an overlapping-stride TMA view; weights are repacked to [Cout][K][8*4 or 4*4].  uda_stem_tc_wgrad accumulates
 * into dw ([Cout][K][K][3] fp32) using dws_scratch (fp32 [Cout][K][32 or 16]). */
int uda_stem_tc_supported(int B, int H, int W, int Cin, int Cout, int K, int stride, int pad);
size_t uda_stem_packed_input_elems(int B, int H, int W);
int uda_stem_pack_input(const float* x_nchw, void* xs, int B, int H, int W, int pad, void* stream);
int uda_stem_pack_weight(const void* w, void* ws, int Cout, int K, void* stream);
int uda_stem_tc_fwd(const void* xs, const void* ws, const float* bias, void* y, double* bn_sums, int B, int H, int W,
                    int Cout, int K, int pad, void* stream);
int uda_stem_tc_fwd_act(const void* xs, const void* ws, const float* bias, void* y, float act_slope, int B, int H, int W,
                        int Cout, int K, int pad, void* stream);
int uda_stem_tc_wgrad(const void* dy, const void* xs, float* dw, float* dws_scratch, int B, int H, int W, int Cout,
                      int K, int pad, void* stream);

/* BatchNorm2d (train: batch statistics; eps 1e-5, momentum 0.1 in the reference's graph).
 * x [M,C] NHWC rows.  uda_bn_stats: writes mean/rstd/scale/shift (float[C]) and updates running statistics
 * when non-NULL; the per-channel finalize runs in the last CTA of the same launch.  uda_bn_stats and uda_bn_bwd
 * share one dedicated 128 KB workspace (C <= 4096) whose first 16 + 2*4096*8 bytes must be ZERO on entry and are
 * left zero on exit (allocate it zeroed once: no per-call memset).  uda_bn_apply: y = act(x*scale+shift (+residual)),
 * slope 1 = identity, 0 = ReLU, 0.2 = LeakyReLU.  uda_bn_bwd:
 * `a` = saved post-activation output (NULL for identity activation); dres (nullable) receives the
 * activation-masked gradient of the residual branch. */
int uda_bn_stats(const void* x, int dtype, long long M, int C, const float* gamma, const float* beta,
                 float* running_mean, float* running_var, float* mean, float* rstd, float* scale, float* shift,
                 float eps, float momentum, void* workspace, void* stream);
int uda_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float* scale, float* shift, int C, float eps, void* stream);
int uda_bn_apply(const void* x, const void* residual, void* y, int dtype, const float* scale, const float* shift,
                 long long M, int C, float slope, void* stream);
/* bn_apply with the batch statistics produced by the convolution epilogue (uda_conv2d_tc_fwd's bn_sums):
 * normalise + activation and, in CTA 0, mean/rstd/scale/shift outputs + running-statistics update. */
int uda_bn_apply_fused(const void* x, const void* residual, void* y, int dtype, const double* sums, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, float* mean, float* rstd,
                       float* scale, float* shift, long long M, int C, float eps, float momentum, float slope,
                       void* stream);
/* The ResNet stem tail bn1 -> relu -> maxpool (torchvision encoder inside smp.Unet, src/models/train.py:572-577) as ONE
 * pass over x: a = act(BN(x)) exactly as uda_bn_apply_fused (a is kept: it is the first skip connection) and
 * (y, idx) = uda_maxpool3x3s2_fwd(a), taken from the rows while they are in shared memory.  bf16 NHWC only, H and W
 * even, C a power of two in [8, 2048], W*C*2 <= ~50 KB; returns UDA_ERR_UNSUPPORTED (nothing launched) otherwise. */
int uda_bn_apply_maxpool_fused(const void* x, void* a, void* y, unsigned char* idx, int dtype, const double* sums,
                               const float* gamma, const float* beta, float* running_mean, float* running_var,
                               float* mean, float* rstd, float* scale, float* shift, int B, int H, int W, int C,
                               float eps, float momentum, float slope, void* stream);
/* a == NULL with scale/shift given: the activation mask is recomputed from x*scale+shift (non-residual layers),
 * which saves reading the saved output in both backward passes. */
int uda_bn_bwd(const void* dy, const void* x, const void* a, int dtype, const float* gamma, const float* mean,
               const float* rstd, const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate,
               float* dgamma, float* dbeta, int param_accumulate, long long M, int C, float slope, void* workspace,
               void* stream);
/* BatchNorm backward with the reduction fused into the producer of dy: uda_conv2d_tc_dgrad_bnstats (the dgrad of the
 * convolution that consumed a = act(BN(z) (+res)), reference: autograd of nn.BatchNorm2d inside smp.Unet) adds
 * sums[c] += sum g, sums[C+c] += sum g*v over its output, g = dy*(a > 0 ? 1 : slope), v = z (v_is_z = 1, z given) or
 * the pre-activation recovered from a (v_is_z = 0; non-residual layers: xhat = (v - beta)/gamma).  sums (double[2C])
 * must be zero before the dgrad.  uda_bn_bwd_apply_fused then finalizes per channel in its prologue (dgamma, dbeta,
 * coefficients) and applies dx (+ the residual-branch gradient) in ONE launch.  bf16, power-of-two C, >= 1 MiB. */
int uda_conv2d_tc_dgrad_bnstats(const void* dy, const void* w_ft, const void* addend, void* dx, int B, int H, int W,
                                int Cin, int Cout, int KH, int KW, int stride, int pad, const void* a, const void* z,
                                float slope, double* sums, void* stream);
int uda_bn_bwd_fused_supported(int dtype, long long M, int C);
int uda_bn_bwd_apply_fused(const void* dy, const void* x, const void* a, int dtype, const double* sums, int v_is_z,
                           const float* gamma, const float* beta, const float* mean, const float* rstd,
                           const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate,
                           float* dgamma, float* dbeta, int param_accumulate, long long M, int C, float slope,
                           void* stream);
int uda_act_bwd(const void* dy, const void* a, void* dx, int dtype, long long n, float slope, void* stream);
int uda_bias_act(const void* x, const float* bias, void* y, int dtype, long long M, int C, float slope, void* stream);
int uda_colsum(const void* x, int dtype, float* out, long long M, int C, float scale, int accumulate,
               void* workspace, void* stream);

int uda_maxpool3x3s2_fwd(const void* x, void* y, unsigned char* idx, int dtype, int B, int H, int W, int C,
                         void* stream);
int uda_maxpool3x3s2_bwd(const void* dy, const unsigned char* idx, const void* addend, void* dx, int dtype, int B,
                         int H, int W, int C, void* stream);
/* nearest x2 upsample of x [B,H/2,W/2,C1] concatenated with skip [B,H,W,C2] (C2 may be 0) */
int uda_upsample2x_concat_fwd(const void* x, const void* skip, void* out, int dtype, int B, int H, int W, int C1,
                              int C2, void* stream);
int uda_upsample2x_concat_bwd(const void* dout, void* dx, void* dskip, int dtype, int B, int H, int W, int C1,
                              int C2, void* stream);
/* discriminator classifier: AdaptiveAvgPool2d(1) + Linear(C,1) + Sigmoid (discriminator.py:37-42) */
int uda_gap_linear_sigmoid_fwd(const void* x, int dtype, const float* w, const float* bias, float* pooled,
                               float* out, int B, long long HW, int C, void* stream);
int uda_gap_linear_sigmoid_bwd(const float* dout, const float* y, const float* pooled, const float* w, float* dw,
                               float* dbias, void* dx, int dtype, int B, long long HW, int C, int accumulate,
                               void* stream);

/* Optimizer: torch.optim.Adam semantics over a flat fp32 buffer
 * (src/models/train.py:461), optional bf16 shadow refresh, optional device-side clip coefficient
 * from uda_grad_clip_coef (clip_grad_norm_, src/models/unsupervised_trainer.py:144; workspace 8 bytes).
 * dev_step (optional, device int): graph-capturable mode - the call increments *dev_step and uses it as the
 * step count for the bias corrections instead of `step`. */
int uda_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                  const float* dev_clip_coef, int* dev_step, void* stream);
int uda_grad_clip_coef(const float* g, long long n, float max_norm, float pre_scale, float* coef, float* norm_out,
                       void* workspace, void* stream);

/* Metrics from the device-resident confusion matrix (one CTA, float64): out[0] in-tree mean IoU (nanmean(diag / (row + col
 * - diag + 1e-7)), src/analysis/metrics.py:29-42), out[1] pixel accuracy, out[2] macro Jaccard with torchmetrics semantics
 * (classes without support ignored, src/models/train.py:209-212,231), out[3] pixels counted, out[4..4+C) per-class in-tree
 * IoU, out[4+C..4+2C) per-class binary Jaccard (0 for 0/0, train.py:236-241). */
int uda_metrics_from_hist(const long long* hist, int C, double* out, void* stream);
/* Sliding-window evaluation glue (BASELINE config 5; window = stride = win, row-major window grid): windows
 * [first, first+count) of a [H,W,3] uint8 tile as normalised fp32 NCHW input ((u8/255 - mean) / std, the reference's
 * ToTensor + Normalize, src/models/predict.py:93-97); the same windows of an int64 / uint8 [H,W] label tile as int64
 * targets; uint8 window masks scattered back into the [H,W] tile mask.  mean3 / std3 are HOST arrays of three floats. */
int uda_gather_windows_u8(const unsigned char* tile_hwc, float* out_nchw, int H, int W, int win, int first, int count,
                          const float* mean3, const float* std3, void* stream);
int uda_gather_label_windows(const void* tile, int dtype, long long* out, int H, int W, int win, int first, int count,
                             void* stream);
int uda_scatter_window_masks(const unsigned char* masks, unsigned char* tile_mask, int H, int W, int win, int first,
                             int count, void* stream);

/* Device-side strong augmentation of a batch (unsupervised fine-tuning; reference: host-side albumentations pipeline,
 * src/models/unsupervised_trainer.py:99-114, src/models/augmentation.py:40-80): one gather pass per view.  table = device
 * array of B rows x 12 floats {m00,m01,m02, m10,m11,m12 (output pixel -> source coordinates, bilinear, BORDER_REFLECT_101),
 * alpha, beta (value * alpha + beta), sigma (additive Gaussian noise), seed, 0, 0}.  images / out: fp32 NCHW, distinct. */
int uda_strong_augment(const float* images, float* out, const float* table, int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UDA_B200_H_ */
