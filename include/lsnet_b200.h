/*
 * liblsnet_sm100.so — C ABI of the B200-native LSNet training hot path.
 *
 * Every entry point: plain pointers + sizes, no torch types; device pointers unless marked host; returns 0 on
 * success, non-zero on failure with a thread-local message from lsnet_last_error().  No allocation inside:
 * callers pass outputs and workspaces.  Kernels are enqueued on `stream` (a cudaStream_t passed as void*).
 * Activations are NHWC ("pixel-major"): a row is one pixel, `ld*` arguments are row pitches in ELEMENTS.
 *
 * Each function names the reference interface it replaces (paths relative to /root/reference/code).  The
 * reference's own native boundary for this path is the pybind module `deform_conv_ext`
 * (mmdet/ops/dcn/src/deform_conv_ext.cpp:227-250) plus `sigmoid_focal_loss_ext`; everything else on the path is
 * PyTorch Python.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 */
#ifndef LSNET_B200_H_
#define LSNET_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element type of activations / packed weights (accumulation is always fp32) */
enum { LSNET_DTYPE_BF16 = 0 };

/* ---- status ------------------------------------------------------------------------------------------------ */
const char* lsnet_last_error(void);             /* message of the last failing call on this thread */
unsigned long long lsnet_launch_count(void);    /* kernels launched by this library since load */
int lsnet_abi_version(void);
int lsnet_require_sm100(void);                  /* 0 iff the current CUDA device is compute capability 10.x */

/* Optional device timing per kernel class (0 tcgen05 GEMM/conv, 1 weight-gradient GEMM, 2 DCN gather, 3 DCN scatter,
 * 4-6 fused DCN forward / weight gradient / backward data, 7 GEMM launches below the FLOP/byte ridge, accounted in bytes):
 * when enabled every launch of the class is bracketed by CUDA events on its own stream; collect() returns the summed
 * elapsed ms, the launch count and the summed algorithmic work (FLOPs for 0/1, bytes for 2/3) since the last reset. */
void lsnet_timing_enable(int on);
int lsnet_timing_collect(int kernel_class, double* total_ms, long long* launches, double* work);
void lsnet_timing_reset(void);

/* ---- tcgen05 GEMM / implicit-GEMM convolution ----------------------------------------------------------------
 * out[M,N] = A[M,K] . Bw[N,K]^T (+ bias[N]) (ReLU);  bf16 operands, fp32 accumulation in TMEM, out bf16 or fp32.
 * Replaces the per-sample `addmm_` of the reference (mmdet/ops/dcn/src/cuda/deform_conv_cuda.cpp:673-678, 890-895,
 * and the bias add :689-691) when A is the DCN column matrix, and 1x1 convolutions (nn.Conv2d in
 * mmdet/models/necks/fpn.py:117-133, mmdet/models/dense_heads/lsnet_head.py:160-184) when A is an NHWC map.
 * Needs N % 16 == 0, K % 8 == 0, 16-byte aligned row pitches. */
int lsnet_gemm_bf16(const void* A, long long lda, const void* Bw, long long ldb, void* out, long long ldc, int M,
                    int N, int K, const float* bias, int relu, int out_fp32, void* stream);
/* lsnet_gemm_bf16 with the full epilogue: v = acc + bias + resid[row,:] (bf16, pitch ldr); ReLU; v = 0 where mask[row,:]
 * (bf16, pitch ldm) <= 0.  The trunk's 1x1 convolutions: conv3 + folded BN + identity add + ReLU of a Bottleneck
 * (mmdet/models/backbones/resnet.py:286-299) is ONE call with resid = the block input. */
int lsnet_gemm_ex_bf16(const void* A, long long lda, const void* Bw, long long ldb, void* out, long long ldc, int M,
                       int N, int K, const float* bias, const void* resid, long long ldr, const void* mask,
                       long long ldm, int relu, int out_fp32, void* stream);
/* Block-diagonal product for grouped weights: out[M, N] (bf16), N = ntiles*64; column tile nt =
 * A[:, (nt % cblks)*64 .. +63] . Bw[nt*64 .. +63, 0..63]^T with A bf16 [M, cblks*64], Bw bf16 [N, 64].  The dCol GEMM of a
 * grouped DCN (per-group `addmm_` loop of deform_conv_cuda.cpp:746-749) without the groups-times redundant dense FLOPs. */
int lsnet_gemm_blockdiag_bf16(const void* A, long long lda, const void* Bw, void* out, long long ldc, int M, int N,
                              int cblks, void* stream);

/* Stride-1 "same" convolution as an implicit GEMM: x NHWC bf16 [B,H,W,C] (pixel pitch ldp), Wt bf16
 * [N, kh*kw*C] (tap-major, channel-minor), out [B*H*W, ldc].  One shifted TMA box per filter tap; zero padding is
 * the TMA out-of-bounds fill.  Replaces the cuDNN nn.Conv2d 3x3 calls of FPN (necks/fpn.py:146-155) and LSHead
 * (dense_heads/lsnet_head.py:167-184, conv_offset of ModulatedDeformConvPack ops/dcn/deform_conv.py:511-519) and,
 * with transposed/flipped weights, their input gradients.  Needs C % 8 == 0 (weights packed with C padded to 64 per tap), N % 16 == 0. */
int lsnet_conv2d_nhwc_bf16(const void* x, int B, int H, int W, int C, long long ldp, const void* Wt, int N, int kh,
                           int kw, int pad_h, int pad_w, int dil_h, int dil_w, void* out, long long ldc,
                           const float* bias, int relu, int out_fp32, void* stream);

/* General strided / phase-decomposed convolution as ONE implicit GEMM (every conv entry point above is a special case):
 *   out[b, ho*osh+ooh, wo*osw+oow, n] = sum_t sum_c x[b, ho*ish+tap_dy[t], wo*isw+tap_dx[t], c] * Wt[n, t*Cpad64 + c]
 *                                       (+ bias[n]) (+ resid[same pixel, n]) (ReLU) (0 where mask[same pixel, n] <= 0)
 * for ho < Ho, wo < Wo; x NHWC bf16 [B,H,W,C] read through a TMA map with element strides (ish, isw) (out-of-range taps
 * are zero fill), out an OH x OW pixel grid with row pitch ldc.  tap_dy / tap_dx: host int arrays, <= 16 taps; Wt holds
 * wt_taps blocks of Cpad64 columns and tap t multiplies block tap_kblk[t] (NULL: block t).
 *  - forward of a stride-s convolution: ish = isw = s, taps (ky*dil - pad, kx*dil - pad), osh = 1 (the trunk's
 *    Bottleneck convs incl. the stride-2 conv2 / downsample, mmdet/models/backbones/resnet.py:165-223, 261-301, and FPN's
 *    stride-2 extra levels, necks/fpn.py:203-211), bias + identity add + ReLU in the epilogue (resnet.py:286-299);
 *  - its input gradient: one call per output phase (osh = s, ooh = phase) over the taps that reach that phase,
 *    Wt = the transposed pack, mask = the ReLU output of the layer below (replaces cuDNN dgrad + a mask pass). */
int lsnet_conv2d_taps_bf16(const void* x, int B, int H, int W, int C, long long ldp, const void* Wt, int N, int ntaps,
                           const int* tap_dy, const int* tap_dx, const int* tap_kblk, int wt_taps, int ish, int isw,
                           int Ho, int Wo, void* out,
                           long long ldc, int osh, int osw, int ooh, int oow, int OH, int OW, const float* bias,
                           const void* resid, long long ldr, const void* mask, long long ldm, int relu, int out_fp32,
                           void* stream);

/* out[M,N] (fp32) += A[P,M]^T . Bm[P,N]  (reduction over the P rows = pixels; split-K, fp32 red.global.add; the
 * caller zero-fills `out`).  Weight-gradient GEMM: replaces `grad_weight[g].addmm_(grad_output, columns^T)`
 * (deform_conv_cuda.cpp:782-787, 1113-1124).  Needs M % 8 == 0, N % 8 == 0. */
int lsnet_gemm_tn_bf16(const void* A, long long lda, const void* Bm, long long ldb, float* out, long long ldc, int P,
                       int M, int N, void* stream);

/* dw[N, kh*kw, C] (fp32) += sum over the Ho x Wo output pixels p of dy[p, n] * x[p*stride - pad + tap*dil, c]: weight
 * gradient of a strided convolution (x through a TMA map with element strides). */
int lsnet_conv2d_wgrad_strided_nhwc_bf16(const void* dy, long long ldy, const void* x, long long ldx, int B, int H, int W,
                                         int C, int Ho, int Wo, int N, int kh, int kw, int stride_h, int stride_w,
                                         int pad_h, int pad_w, int dil_h, int dil_w, float* dw, void* stream);
/* dw[N, kh*kw, C] (fp32) += sum_pixels dy[p, n] * x[p + tap, c]: weight gradient of lsnet_conv2d_nhwc_bf16. */
int lsnet_conv2d_wgrad_nhwc_bf16(const void* dy, long long ldy, const void* x, long long ldx, int B, int H, int W,
                                 int C, int N, int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, float* dw,
                                 void* stream);

/* ---- GroupNorm (+ residual add, + ReLU) over pixel-major bf16 maps: nn.GroupNorm(32, 256) sites of FPN / LSHead
 * (mmdet/models/necks/fpn.py:117-133; mmdet/models/dense_heads/lsnet_head.py:97-113, 700-708, 1843).
 * y = relu?(GN(x (+ x2))).  stats / ws_bstats: workspaces of 3*B*G + 1 doubles (2*B*G fp64 sums, a ticket counter, then
 * the fp32 (mean, rstd) / (s1, s2) table written by the last CTA of the statistics kernel); stats is written by fwd
 * and read by bwd.  (C/G) % 8 == 0. */
int lsnet_groupnorm_fwd(const void* x, long long ldx, const void* x2, long long ldx2, int B, int HW, int C, int G,
                        const float* gamma, const float* beta, float eps, int relu, double* stats, void* y,
                        long long ldy, void* stream);
/* lsnet_groupnorm_fwd whose statistics (the 2*B*G sums at the start of `stats`) were already accumulated by the producer's
 * epilogue (lsnet_dcn_forward_gn): finalises them into the (mean, rstd) table and applies. */
int lsnet_groupnorm_fwd_pre(const void* x, long long ldx, int B, int HW, int C, int G, const float* gamma, const float* beta,
                            float eps, int relu, double* stats, void* y, long long ldy, void* stream);
int lsnet_groupnorm_bwd(const void* x, long long ldx, const void* x2, long long ldx2, const void* dy, long long lddy,
                        int B, int HW, int C, int G, const float* gamma, const float* beta, float eps, int relu,
                        const double* stats, double* ws_bstats, void* dx, long long lddx, float* dgamma, float* dbeta,
                        void* stream);
/* _acc: dgamma / dbeta are ADDED to their targets (the parameters' gradient memory: no memset, no accumulate kernel).
 * dx_colsum (optional, fp32 [C]): the per-channel sum of dx is ADDED there -- the bias gradient of a conv / DCN whose
 * output is this norm's input (replaces that layer's own column-sum pass over dx). */
int lsnet_groupnorm_bwd_acc(const void* x, long long ldx, const void* x2, long long ldx2, const void* dy, long long lddy,
                            int B, int HW, int C, int G, const float* gamma, const float* beta, float eps, int relu,
                            const double* stats, double* ws_bstats, void* dx, long long lddx, float* dgamma,
                            float* dbeta, float* dx_colsum, void* stream);

/* LSHead element-wise glue.  pred_reg: o fp32 [P, ldo] = output of pts_*_init_out (n_out channels); sp[P, ldsp] =
 * softplus(o[:, :n_sp]) (nn.Softplus defaults, lsnet_head.py:96); off[P, ldoff] = the n_off DCN sampling offsets of
 * LSHead.get_pred_reg (lsnet_head.py:372-400) minus dcn_base_offset: entry j is the signed maximum of the slot pair
 * (sp[src[j]], sp[src[j]+1]) (mode 0; the '-' slot wins ties and is negated) or the raw channel o[src[j]] (mode 1, the 8
 * free offsets of task 'bbox').  src / mode / base are HOST arrays of n_off entries.  Backward: go[P, ldgo] (n_out
 * channels) from gsp (may be NULL) and goff (may be NULL), the offset path scaled by gradient_mul
 * ((1 - gm) * reg.detach() + gm * reg, lsnet_head.py:585-587).
 * add_softplus: out = softplus(t + s) (gy NULL) or gy * softplus'(t + s) (refine = softplus(raw + init.detach()),
 * lsnet_head.py:735-755). */
int lsnet_pred_reg_fwd(const float* o, long long ldo, long long P, int n_sp, int n_out, int n_off, const int* src,
                       const int* mode, const float* base, float* sp, long long ldsp, float* off, long long ldoff,
                       void* stream);
int lsnet_pred_reg_bwd(const float* o, long long ldo, long long P, int n_sp, int n_out, int n_off, const int* src,
                       const int* mode, const float* base, const float* gsp, long long ldgsp, const float* goff,
                       long long ldgoff, float gradient_mul, float* go, long long ldgo, void* stream);
int lsnet_add_softplus(const float* t, long long ldt, const float* s, long long lds, const float* gy, long long ldgy,
                       long long P, int C, float* out, long long ldout, void* stream);

/* One-pass optimizer step over flat fp32 buffers: L2 clip at max_norm (grad_norm: device scalar holding |g|; NULL or
 * max_norm <= 0 disables the clip), then SGD with momentum and weight decay (dampening 0, no nesterov).  Replaces
 * clip_grad_norm_ + torch.optim.SGD.step of mmcv's OptimizerHook.after_train_iter
 * (mmcv/mmcv/runner/hooks/optimizer.py:19-28) for the GraphTrainer's flat storage. */
int lsnet_sgd_momentum_step(float* params, const float* grads, float* momentum_buf, long long n,
                            const float* grad_norm, float max_norm, float lr, float momentum, float weight_decay,
                            void* stream);

/* Backward-pass gradient staging: dY (fp32 if gy_fp32 else bf16; P rows of C channels, pitch ldg) -> bf16 rows of Cpad
 * channels (zero padded) in `out`, optionally masked by relu_out > 0 (bf16, the forward output of a conv+ReLU), and the
 * per-channel sum of the staged values into colsum[C] (fp32; the bias gradient, replacing
 * `grad_bias.addmv_(grad_output, ones)` of deform_conv_cuda.cpp:788-794).  relu_out / colsum may be NULL. */
int lsnet_grad_prep(const void* gy, int gy_fp32, long long ldg, const void* relu_out, long long ldo, long long P, int C,
                    int Cpad, void* out, long long ldout, float* colsum, void* stream);
/* _acc: the column sums are ADDED to colsum (the bias parameter's gradient memory). */
int lsnet_grad_prep_acc(const void* gy, int gy_fp32, long long ldg, const void* relu_out, long long ldo, long long P,
                        int C, int Cpad, void* out, long long ldout, float* colsum, void* stream);

/* Frozen-statistics BatchNorm folded into the preceding conv (ResNet trunk with norm_eval=True,
 * mmdet/models/backbones/resnet.py:636-646): Wb[o] = W[o]*s[o] (bf16, OHWI order), bias[o] = beta[o] - mean[o]*s[o],
 * s = gamma/sqrt(var+eps); W is OIHW fp32 with I input channels and KK = kh*kw taps.  bwd: gW (fp32 OIHW), ggamma, gbeta
 * from gWb (bf16 OHWI) and gbias (fp32, may be NULL). */
int lsnet_bn_fold_fwd(const float* W, const float* gamma, const float* beta, const float* mean, const float* var,
                      float eps, int O, int I, int KK, void* Wb, float* bias, void* stream);
int lsnet_bn_fold_bwd(const void* gWb, const float* gbias, const float* W, const float* gamma, const float* mean,
                      const float* var, float eps, int O, int I, int KK, float* gW, float* ggamma, float* gbeta,
                      void* stream);
/* The same fold for the library's own trunk convolutions: W element (o,i,k) at o*I*KK + i*si + k*sk (OIHW: si=KK, sk=1;
 * tap-major: si=1, sk=I).  fwd writes both GEMM operand packs -- Wb bf16 [O, KK*I] (forward / weight-gradient order) and
 * Wt bf16 [I, KK*O] (B operand of the input-gradient GEMM).  bwd takes gWb fp32 [O, KK*I] as lsnet_conv2d_wgrad_* leaves
 * it; accumulate != 0 adds gW / ggamma / gbeta into the parameters' gradient memory. */
int lsnet_bn_fold2_fwd(const float* W, long long si, long long sk, const float* gamma, const float* beta,
                       const float* mean, const float* var, float eps, int O, int I, int KK, void* Wb, void* Wt,
                       float* bias, void* stream);
int lsnet_bn_fold2_bwd(const float* gWb, const float* gbias, const float* W, long long si, long long sk,
                       const float* gamma, const float* mean, const float* var, float eps, int O, int I, int KK,
                       float* gW, float* ggamma, float* gbeta, int accumulate, void* stream);

/* ResNet stem (mmdet/models/backbones/resnet.py:509-520, 619-623: conv1 7x7/2 -> norm1 (frozen) -> relu -> maxpool 3x3/2).
 * x: [B,3,H,W] image, fp32 (x_bf16 = 0) or bf16, ANY element strides (sb, sc, sh, sw) -- NCHW or NHWC; Wp bf16 [64, 192]:
 * column ky*24 + kx*3 + ch holds the BN-folded weight of tap (ky,kx), channel ch, zeros elsewhere; bias fp32 [64] = the BN
 * shift; out bf16 NHWC [B, (H-1)/2+1, (W-1)/2+1, 64] = relu(conv + bias).  Implicit GEMM on tcgen05 whose A tile is built in
 * shared memory from the input window. */
int lsnet_stem_conv7x7s2_bf16(const void* x, int x_bf16, long long sb, long long sc, long long sh, long long sw, int B,
                              int H, int W, const void* Wp, const float* bias, void* out, void* stream);
/* y[B, (H-1)/2+1, (W-1)/2+1, C] = 3x3 / stride 2 / pad 1 max-pool of x NHWC bf16 (C % 8 == 0). */
int lsnet_maxpool3x3s2_nhwc_bf16(const void* x, int B, int H, int W, int C, void* y, void* stream);

/* FPN top-down pathway (mmdet/models/necks/fpn.py:180-192): out = fine + nearest-upsample(coarse) to fine's size
 * (PyTorch 'nearest': source index min(floor(dst * in/out), in-1)); _bwd: gc = sum of g over the fine pixels of every
 * coarse cell (the gradient w.r.t. `fine` is g itself).  Pixel-major bf16, C % 8 == 0. */
int lsnet_upsample_add_nhwc_bf16(const void* fine, long long ldf, const void* coarse, long long ldc, int B, int Hf, int Wf,
                                 int Hc, int Wc, int C, void* out, long long ldo, void* stream);
int lsnet_upsample_add_bwd_nhwc_bf16(const void* g, long long ldg, int B, int Hf, int Wf, int Hc, int Wc, int C, void* gc,
                                     long long ldgc, void* stream);

/* ---- input side -------------------------------------------------------------------------------------
 * Normalize -> Pad(size_divisor) -> DefaultFormatBundle -> collate of the reference's train pipeline in one pass
 * (mmdet/datasets/pipelines/transforms.py:463-570, formating.py:209-215, mmcv/mmcv/image/photometric.py:21-41,
 * mmcv/mmcv/parallel/collate.py:39-60).  src: uint8 [B, H, W, 3] (decoded / resized / flipped BGR images, top-left
 * aligned on a common canvas, W % 4 == 0); hw: device int32 [B][2] = each image's own (height, width);
 * dst: fp32 [B, H, W, 3] (the NHWC memory of a channels_last [B, 3, H, W] tensor, what lsnet_stem_conv reads):
 * dst[b,y,x,c] = float((double(src[b,y,x, to_rgb ? 2-c : c]) - mean[c]) * stdinv[c]) inside the image, 0 in the padding
 * (double arithmetic, one rounding: bit-identical to the reference's cv2.subtract / cv2.multiply). */
int lsnet_image_prep_u8(const void* src_u8, const int* hw, int B, int H, int W, double mean0, double mean1, double mean2,
                        double stdinv0, double stdinv1, double stdinv2, int to_rgb, void* dst_f32, void* stream);

/* ---- deformable convolution sampling ---------------------------------------------------------------------------
 * One family for DCNv1 (mask NULL), DCNv2 (mask) and LSNet's pyramid DCN (scale_h/scale_w, input extent (H,W)
 * decoupled from the sampling grid (Ho,Wo)).  x: NHWC bf16; offset: fp32 [B*Ho*Wo, ldo], channel
 * g*2*kh*kw + 2*k + {0:dy, 1:dx}; mask: fp32 [B*Ho*Wo, ldm], channel g*kh*kw + k; col: bf16 [B*Ho*Wo, kh*kw*C].
 * mask_logits = 1: `mask` holds the raw conv_offset outputs and the kernels apply the sigmoid of
 * ModulatedDeformConvPack.forward (deform_conv.py:528-531) themselves (the adjoint then returns dmask w.r.t. the
 * logits), so offset and mask can be two column ranges of ONE conv_offset output and no element-wise kernel runs.
 * Replaces deformable_im2col / pyramid_deformable_im2col / modulated_deformable_im2col_cuda
 * (mmdet/ops/dcn/src/cuda/deform_conv_cuda_kernel.cu:190-332, 647-680, 847-910, 1046-1076), i.e. the sampling half
 * of deform_conv_forward / pyramid_deform_conv_forward / modulated_deform_conv_forward
 * (ops/dcn/src/deform_conv_ext.cpp:74-147). */
int lsnet_dcn_im2col_bf16(const void* x, int B, int H, int W, int C, long long ldx, const float* offset,
                          long long ldo, const float* mask, long long ldm, int Ho, int Wo, int kh, int kw,
                          int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, float scale_h,
                          float scale_w, int deformable_groups, void* col, long long ldcol, int mask_logits,
                          void* stream);

/* Adjoint: gcol bf16 [B*Ho*Wo, kh*kw*C] (= dY . W) -> dx NHWC, ACCUMULATED (caller zero-fills; may be NULL): bf16 with
 * 16-byte packed-bf16 vector reds (dx_fp32 = 0, the training default) or fp32 with v4.f32 reds (dx_fp32 = 1),
 * doffset fp32 [B*Ho*Wo, lddo], dmask fp32 [B*Ho*Wo, lddm] (NULL without mask).  Replaces *_col2im and
 * *_col2im_coord (deform_conv_cuda_kernel.cu:333-448, 486-615, 912-1044), the sampling half of
 * deform_conv_backward_input / pyramid_deform_conv_backward_input / modulated_deform_conv_backward
 * (deform_conv_ext.cpp:92-224). */
int lsnet_dcn_col2im_bf16(const void* gcol, long long ldcol, const void* x, int B, int H, int W, int C, long long ldx,
                          const float* offset, long long ldo, const float* mask, long long ldm, int Ho, int Wo,
                          int kh, int kw, int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                          float scale_h, float scale_w, int deformable_groups, void* dx, long long lddx,
                          int dx_fp32, float* doffset, long long lddo, float* dmask, long long lddm, int mask_logits,
                          void* stream);

/* ---- deformable convolution as whole operators ---------------------------------------------------------------------
 * One call per pybind function of the reference's `deform_conv_ext` (mmdet/ops/dcn/src/deform_conv_ext.cpp:227-250):
 *   lsnet_dcn_forward          deform_conv_forward :74-90 / modulated_deform_conv_forward :129-147 /
 *                              pyramid_deform_conv_forward (deform_conv_cuda.cpp:811-919)
 *   lsnet_dcn_backward_data    deform_conv_backward_input :92-109 / pyramid_..._backward_input (deform_conv_cuda.cpp:921-1036)
 *                              and the input/offset/mask half of modulated_deform_conv_backward :149-224
 *   lsnet_dcn_backward_weight  deform_conv_backward_parameters :111-127 / pyramid_..._backward_parameters
 *                              (deform_conv_cuda.cpp:1038-1154) and the weight half of modulated_deform_conv_backward
 * The descriptor carries the arguments those functions share (kernel / stride / pad / dilation / groups /
 * deformable_groups, and LSNet's pyramid scales).  `mask` NULL = DCNv1 / pyramid; mask_logits = 1: `mask` holds the raw
 * conv_offset outputs (sigmoid applied inside, dmask returned w.r.t. the logits).  Layouts as for lsnet_dcn_im2col_bf16.
 * forward: the FUSED kernel (dcn_fused.cu: bilinear gather -> SWIZZLE_128B shared-memory A tile -> tcgen05.mma, no column
 * matrix in HBM) runs when deformable_groups == 1, kh*kw <= 9, C % 64 == 0 and N <= 256; other shapes take gather -> GEMM
 * through `workspace`.  *_workspace_size returns the scratch bytes the matching call needs (0 = none). */
typedef struct lsnet_dcn_desc {
  int B, H, W, C;              /* sampled map x: NHWC, C channels */
  long long ldx;               /* pixel pitch of x in elements */
  int Ho, Wo;                  /* sampling / output grid (= offset grid) */
  int kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  float scale_h, scale_w;      /* pyramid DCN: base grid scale (1 for DCNv1 / DCNv2) */
  int groups;                  /* weight groups; > 1 runs the block-diagonal kernels when lsnet_dcn_grouped_supported() */
  int deformable_groups;
  int mask_logits;
  int dtype;                   /* LSNET_DTYPE_BF16 */
} lsnet_dcn_desc;

/* A/B switch for tests and measurements: 0 = column-matrix path for every shape, 1 = fused kernels wherever the shape is
 * supported, 2 = automatic (fused where it measured faster: patches >= 2 waves of the SMs).  Default: env LSNET_DCN_FUSED,
 * else 2.  Not thread-safe; call between steps. */
void lsnet_dcn_fused_enable(int on);
/* 1: the weight gradient of lsnet_dcn_backward_weight is reduced in a fixed order (per-split partials in the workspace +
 * a second pass) instead of fp32 atomics whose order varies from run to run (default: env LSNET_DETERMINISTIC, else 0). */
void lsnet_set_deterministic(int on);
/* Grouped weights (ResNeXt conv2 sites, mmdet/models/backbones/resnext.py:50-74: groups = 64, C = N = 512/1024/2048) run on
 * 64-channel blocks that hold whole groups -- 1 when the descriptor qualifies (C == N, C % 256 == 0, 64 % (C/groups) == 0,
 * deformable_groups == 1).  Operand layouts then are:
 *   forward          Wp  bf16 [C, kh*kw*64]: row o, column tap*64 + j = weight of out channel o for input channel
 *                        (o/64)*64 + j (zero when that channel is outside o's group)
 *   backward_data    Wt  bf16 [kh*kw*C, 64]: row tap*C + blk*64 + j, column o' = weight of out channel blk*64 + o' for
 *                        input channel blk*64 + j (zero outside the group)
 *   backward_weight  dW  fp32 [C, kh*kw*256] (+=): row o, column tap*256 + c' = gradient for input channel (o/256)*256 + c'
 *                        (the caller keeps the entries of o's own group)
 * Otherwise callers expand the weight to a dense block-diagonal pack and pass groups = 1. */
int lsnet_dcn_grouped_supported(const lsnet_dcn_desc* d, int N);
size_t lsnet_dcn_forward_workspace_size(const lsnet_dcn_desc* d, int N);
/* out[B*Ho*Wo, ldc] (bf16, or fp32 if out_fp32) = DCN(x; offset, mask) . Wp^T (+ bias) (ReLU).  Wp: bf16 [N, kh*kw*C]
 * (tap-major, channel-minor), N % 16 == 0 (rows beyond the real Cout are zero).  col_out (optional, bf16
 * [B*Ho*Wo, kh*kw*C]): also emit the column matrix (operand of the unfused weight gradient). */
int lsnet_dcn_forward(const lsnet_dcn_desc* d, const void* x, const float* offset, long long ldo, const float* mask,
                      long long ldm, const void* Wp, int N, const float* bias, int relu, void* out, long long ldc,
                      int out_fp32, void* col_out, void* workspace, size_t workspace_bytes, void* stream);

/* lsnet_dcn_forward that also accumulates the GroupNorm statistics of its output (before ReLU) in the GEMM epilogue
 * (DCNConvModule = conv -> GroupNorm -> ReLU, mmdet/models/dense_heads/lsnet_head.py:1830-1849): gn_sums is
 * the first 2*B*G doubles of a lsnet_groupnorm_fwd workspace, ZEROED by the caller; the kernel adds (sum x, sum x^2) per
 * (image, group).  lsnet_groupnorm_fwd_pre then normalises without its own statistics pass.  NULL = plain forward. */
int lsnet_dcn_forward_gn(const lsnet_dcn_desc* d, const void* x, const float* offset, long long ldo, const float* mask,
                         long long ldm, const void* Wp, int N, const float* bias, int relu, void* out, long long ldc,
                         int out_fp32, void* col_out, void* workspace, size_t workspace_bytes, double* gn_sums,
                         int gn_groups, void* stream);

size_t lsnet_dcn_backward_data_workspace_size(const lsnet_dcn_desc* d, int N);
/* dy: bf16 [B*Ho*Wo, ldy] (N channels, N % 8 == 0); Wt: bf16 [kh*kw*C, N] (= Wp^T).  dx (NHWC, ACCUMULATED: the caller
 * zero-fills; bf16, or fp32 if dx_fp32; may be NULL), doffset fp32 [.., lddo], dmask fp32 [.., lddm] (NULL without mask). */
int lsnet_dcn_backward_data(const lsnet_dcn_desc* d, const void* dy, long long ldy, int N, const void* Wt, const void* x,
                            const float* offset, long long ldo, const float* mask, long long ldm, void* dx,
                            long long lddx, int dx_fp32, float* doffset, long long lddo, float* dmask, long long lddm,
                            void* workspace, size_t workspace_bytes, void* stream);

size_t lsnet_dcn_backward_weight_workspace_size(const lsnet_dcn_desc* d, int N, int have_col);
/* dW[N, lddw] (fp32, tap-major [N, kh*kw*C]) += dy^T . columns.  col_saved: the column matrix emitted by the forward
 * (col_out), or NULL to re-sample x -- inside the FUSED weight-gradient kernel (gather warps write the MN-major B tiles of
 * a split-K tcgen05 GEMM; needs deformable_groups == 1, kh*kw <= 9, C % 256 == 0), else through `workspace`. */
int lsnet_dcn_backward_weight(const lsnet_dcn_desc* d, const void* dy, long long ldy, int N, const void* x,
                              const float* offset, long long ldo, const float* mask, long long ldm, const void* col_saved,
                              float* dW, long long lddw, void* workspace, size_t workspace_bytes, void* stream);

/* ---- cross-IOU loss ----------------------------------------------------------------------------------------------
 * loss_type: 0 bbox, 1 polygon, 2 keypoint.  Dense form = CrossIOULoss.forward / cross_iou_loss
 * (mmdet/models/losses/cross_iou_loss.py:61-172): pred/target [N,D] fp32, pos_inds [N,D] u8, weight_row [N]
 * (= weight.mean(-1)), anchor_pts [N,2], bbox_gt [N,4], vs [N,L].  row_loss[n] = weight_row[n] * loss_n. */
int lsnet_cross_iou_fwd(int loss_type, const float* pred, const float* target, const unsigned char* pos_inds,
                        const float* weight_row, const float* anchor_pts, const float* bbox_gt, const float* vs, int N,
                        int D, int L, float eps, float alpha, int stride, float* row_loss, void* stream);
/* dpred[n,d] = (*scale) * weight_row[n] * d loss_n / d pred[n,d]   (scale: device scalar) */
int lsnet_cross_iou_bwd(int loss_type, const float* pred, const float* target, const unsigned char* pos_inds,
                        const float* weight_row, const float* anchor_pts, const float* bbox_gt, const float* vs, int N,
                        int D, int L, float eps, float alpha, int stride, const float* scale, float* dpred,
                        void* stream);
/* Fused per-level form used by LSHead.loss (dense_heads/lsnet_head.py:1064-1102, 402-454): rows are the (b,h,w)
 * pixels of an NHWC prediction map; targets are rebuilt from assign[b, level_off + pix] (index into the per-image GT
 * tables gt_pts [B,Gmax,2*NP], gt_bbox [B,Gmax,4], gt_vs [B,Gmax,L]).  backward == 0: row_loss; else dpred. */
int lsnet_cross_iou_level(int loss_type, int backward, const float* pred, long long ldp, int D, const int* assign,
                          long long assign_ld, int level_off, int B, int Hl, int Wl, float stride, float base_scale,
                          const float* gt_pts, const float* gt_bbox, const float* gt_vs, int Gmax, int NP, int L,
                          float eps, float alpha, int pstride, float* row_loss, const float* scale, float* dpred,
                          void* stream);
/* LSHead.get_bbox_gt_reg / get_poly_gt_reg (lsnet_head.py:402-454), dense rows. */
int lsnet_directional_targets(const float* gt_rows, const float* anchor_pts, const float* weight_row, int N, int NP,
                              float* target, unsigned char* pos_inds, void* stream);

/* ---- sigmoid focal loss: SigmoidFocalLossForward/Backward (mmdet/ops/sigmoid_focal_loss/src/cuda/
 * sigmoid_focal_loss_cuda.cu:23-97, sigmoid_focal_loss_ext.cpp).  labels int32, background = C. ------------------ */
int lsnet_focal_partial_count(long long N, int C);   /* host: number of partial sums lsnet_focal_fwd writes */
int lsnet_focal_fwd(const float* logits, long long ldl, const int* labels, const float* weight, long long N, int C,
                    float gamma, float alpha, float* partial, void* stream);
int lsnet_focal_bwd(const float* logits, long long ldl, const int* labels, const float* weight, long long N, int C,
                    float gamma, float alpha, const float* scale, float* dlogits, long long ldd, void* stream);

/* ---- landmark-target assignment (bit-exact integers on tie-free inputs) -----------------------------------------
 * Pyramid description (host arrays): level_h/level_w/level_stride [num_levels]; points of level l are
 * (w*stride, h*stride, stride) (mmdet/core/anchor/point_generator.py:17-25).  valid_hw: device int [B,num_levels,2].
 * gt_bbox [B,Gmax,4], gt_count [B] (device).  assign: int32 [B, sum_l H_l*W_l], -1 = background, else GT index.  */
/* CentroidAssigner.assign, iou_type='center', pos_num=1 (mmdet/core/bbox/assigners/centroid_assigner.py:26-93). */
int lsnet_centroid_assign(int num_levels, const int* level_h, const int* level_w, const float* level_stride,
                          const int* valid_hw, const float* gt_bbox, const int* gt_count, int B, int Gmax, float scale,
                          float* ws_best_d, int* ws_best_i, int* assign, void* stream);
/* ATSSAssigner.assign (mmdet/core/bbox/assigners/atss_assigner.py:29-164); boxes [B,total,4] predicted init boxes;
 * ws_keys: u64 [B,total] workspace; max_overlaps may be NULL. */
int lsnet_atss_assign(int num_levels, const int* level_h, const int* level_w, const float* level_stride,
                      const int* valid_hw, const float* boxes, const float* gt_bbox, const int* gt_count, int B,
                      int Gmax, int topk, unsigned long long* ws_keys, int* assign, float* max_overlaps, void* stream);
/* LSHead._target_single label scatter (lsnet_head.py:834-898): labels int32 (background = num_classes),
 * label_weights, per-image positive counts num_pos [B]. */
int lsnet_assign_targets(int num_levels, const int* level_h, const int* level_w, const float* level_stride,
                         const int* valid_hw, const int* assign, const int* gt_labels, int B, int Gmax,
                         int num_classes, int* labels, float* label_weights, int* num_pos, void* stream);
/* extreme_points2bbox / vectors2bbox + centre shift (lsnet_head.py:321-370, 1333-1361). */
int lsnet_pred_boxes(const float* pred, long long ldp, int NP, int polygon, int B, int Hl, int Wl, float stride,
                     int level_off, int total_points, float* boxes, void* stream);

/* ---- inference decode: greedy NMS (mmdet/ops/nms/nms_wrapper.py:7-157, src/cuda/nms_kernel.cu; called from
 * multiclass_nms_lsvr, mmdet/core/post_processing/bbox_nms.py:60-99).  boxes: fp32 [n,4] (x1,y1,x2,y2), 16-byte aligned,
 * sorted by DESCENDING score; box j is dropped when an earlier kept box overlaps it with IoU > iou_thr (areas without +1).
 * keep[n]: indices of the kept boxes in order, *num_keep their count (both device).  No host synchronisation. */
size_t lsnet_nms_workspace_size(int n);
int lsnet_nms(const float* boxes, int n, float iou_thr, void* workspace, size_t workspace_bytes, int* keep, int* num_keep,
              void* stream);

/* ---- host-side parity hooks (CPU, no device): the row arithmetic the loss kernels run ----------------------- */
float lsnet_host_cross_iou_row(int loss_type, const float* pred, const float* target, const unsigned char* pos_inds,
                               int D, const float* anchor, const float* bbox_gt, const float* vs, float eps,
                               float alpha, int stride, float* grad);
float lsnet_host_focal_elem(float x, int t, int d, float gamma, float alpha, float* grad);
void lsnet_host_directional_target_row(const float* gt, int NP, const float* anchor, int positive, float* target,
                                       unsigned char* sel);

#ifdef __cplusplus
}
#endif
#endif /* LSNET_B200_H_ */
