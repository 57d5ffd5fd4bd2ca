/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked or imported by the product
 * package (lsnet_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Plain-C CPU restatement of the reference's deformable-convolution family
 * (DCNv1 `deform_conv`, DCNv2 `modulated_deform_conv`, LSNet's
 * `pyramid_deform_conv`).  The reference only ships CUDA kernels for these
 * (mmdet/ops/dcn/src/deform_conv_ext.cpp:89-224 -> "not implemented on CPU"),
 * so this file restates their arithmetic, citing the reference lines it
 * follows.  Layouts are the reference's: NCHW activations, offsets
 * (B, dg*2*kh*kw, Ho, Wo) interleaved (dy,dx) per tap, masks (B, dg*kh*kw, Ho, Wo),
 * column matrix [C*kh*kw][B][Ho][Wo]  (deform_conv_cuda_kernel.cu:211,223-224).
 *
 * One code path serves all three ops:
 *   mask == NULL            -> DCNv1 / pyramid   (…kernel.cu:190-297)
 *   mask != NULL            -> DCNv2             (…kernel.cu:847-910)
 *   scale_h/scale_w != 1,
 *   (H,W) != input of (Ho,Wo)-> pyramid          (…kernel.cu:281-282)
 * Parity status: pinned on the GPU box against the reference's own CUDA
 * extension compiled into oracle/_ref (tests/test_dcn_vs_reference_cuda.py) and,
 * everywhere, against torchvision.ops.deform_conv2d (independent implementation).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define DEFINE_DCN(T, SUF)                                                                    \
  /* bilinear sample with per-corner zeroing: …kernel.cu:84-115 (and :744-774) */           \
  static T bilinear_##SUF(const T *im, int H, int W, T h, T w) {                              \
    int h_low = (int)floor((double)h), w_low = (int)floor((double)w);                        \
    int h_high = h_low + 1, w_high = w_low + 1;                                               \
    T lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;                               \
    T v1 = 0, v2 = 0, v3 = 0, v4 = 0;                                                         \
    if (h_low >= 0 && w_low >= 0) v1 = im[h_low * W + w_low];                                 \
    if (h_low >= 0 && w_high <= W - 1) v2 = im[h_low * W + w_high];                           \
    if (h_high <= H - 1 && w_low >= 0) v3 = im[h_high * W + w_low];                           \
    if (h_high <= H - 1 && w_high <= W - 1) v4 = im[h_high * W + w_high];                     \
    T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;                                 \
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;                                             \
  }                                                                                           \
  /* d(sample)/d(h) (dir 0) or d(w) (dir 1): …kernel.cu:145-188 */                           \
  static T coord_weight_##SUF(const T *im, int H, int W, T h, T w, int dir) {                 \
    if (h <= -1 || h >= H || w <= -1 || w >= W) return 0;                                     \
    int h_low = (int)floor((double)h), w_low = (int)floor((double)w);                        \
    int h_high = h_low + 1, w_high = w_low + 1;                                               \
    T wt = 0;                                                                                 \
    if (dir == 0) {                                                                           \
      if (h_low >= 0 && w_low >= 0) wt += -1 * (w_low + 1 - w) * im[h_low * W + w_low];       \
      if (h_low >= 0 && w_high <= W - 1) wt += -1 * (w - w_low) * im[h_low * W + w_high];     \
      if (h_high <= H - 1 && w_low >= 0) wt += (w_low + 1 - w) * im[h_high * W + w_low];      \
      if (h_high <= H - 1 && w_high <= W - 1) wt += (w - w_low) * im[h_high * W + w_high];    \
    } else {                                                                                  \
      if (h_low >= 0 && w_low >= 0) wt += -1 * (h_low + 1 - h) * im[h_low * W + w_low];       \
      if (h_low >= 0 && w_high <= W - 1) wt += (h_low + 1 - h) * im[h_low * W + w_high];      \
      if (h_high <= H - 1 && w_low >= 0) wt += -1 * (h - h_low) * im[h_high * W + w_low];     \
      if (h_high <= H - 1 && w_high <= W - 1) wt += (h - h_low) * im[h_high * W + w_high];    \
    }                                                                                         \
    return wt;                                                                                \
  }                                                                                           \
  /* sampling position of tap (i,j) at output (ho,wo): …kernel.cu:226-227 (v1),              \
     :281-282 (pyramid: the base grid is scaled, the offset is not), :884-885 (v2) */         \
  static void sample_pos_##SUF(const T *off_b, int k, int kk, int Ho, int Wo, int ho, int wo, \
                               int i, int j, int sh, int sw, int ph, int pw, int dh, int dw,  \
                               float scale_h, float scale_w, T *h, T *w) {                    \
    (void)kk;                                                                                 \
    T oh = off_b[((size_t)(2 * k) * Ho + ho) * Wo + wo];                                      \
    T ow = off_b[((size_t)(2 * k + 1) * Ho + ho) * Wo + wo];                                  \
    int h_in = ho * sh - ph, w_in = wo * sw - pw;                                             \
    *h = (T)((float)(h_in + i * dh) * scale_h) + oh;                                          \
    *w = (T)((float)(w_in + j * dw) * scale_w) + ow;                                          \
  }                                                                                           \
  /* im2col: …kernel.cu:190-243 / 245-297 / 847-910.  col[(c*kk+k)][b][ho][wo] */            \
  void dcn_im2col_##SUF(const T *x, const T *offset, const T *mask, int B, int C, int H,      \
                        int W, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph,        \
                        int pw, int dh, int dw, float scale_h, float scale_w, int dg,         \
                        T *col) {                                                             \
    int kk = kh * kw, cpg = C / dg;                                                           \
    _Pragma("omp parallel for schedule(static)")                                              \
    for (int c = 0; c < C; ++c) {                                                             \
      int g = c / cpg;                                                                        \
      for (int b = 0; b < B; ++b) {                                                           \
        const T *im = x + ((size_t)b * C + c) * H * W;                                        \
        const T *off_b = offset + ((size_t)b * dg + g) * 2 * kk * Ho * Wo;                    \
        const T *msk_b = mask ? mask + ((size_t)b * dg + g) * kk * Ho * Wo : NULL;            \
        for (int i = 0; i < kh; ++i)                                                          \
          for (int j = 0; j < kw; ++j) {                                                      \
            int k = i * kw + j;                                                               \
            T *dst = col + (((size_t)(c * kk + k) * B + b) * Ho) * Wo;                        \
            for (int ho = 0; ho < Ho; ++ho)                                                   \
              for (int wo = 0; wo < Wo; ++wo) {                                               \
                T h, w, val = 0;                                                              \
                sample_pos_##SUF(off_b, k, kk, Ho, Wo, ho, wo, i, j, sh, sw, ph, pw, dh, dw,  \
                                 scale_h, scale_w, &h, &w);                                   \
                if (h > -1 && w > -1 && h < H && w < W) val = bilinear_##SUF(im, H, W, h, w); \
                if (msk_b) val *= msk_b[((size_t)k * Ho + ho) * Wo + wo];                     \
                dst[(size_t)ho * Wo + wo] = val;                                              \
              }                                                                               \
          }                                                                                   \
      }                                                                                       \
    }                                                                                         \
  }                                                                                           \
  /* col2im: scatter grad_col to the <=4 in-range corners with the bilinear weights.          \
     …kernel.cu:333-389 / 391-448 / 912-970 reach the same four corners through an (int)     \
     truncation + 5x5 search with |d|<1 and get_gradient_weight (:117-143). grad_im +=. */    \
  void dcn_col2im_##SUF(const T *gcol, const T *offset, const T *mask, int B, int C, int H,   \
                        int W, int Ho, int Wo, int kh, int kw, int sh, int sw, int ph,        \
                        int pw, int dh, int dw, float scale_h, float scale_w, int dg,         \
                        T *grad_im) {                                                         \
    int kk = kh * kw, cpg = C / dg;                                                           \
    _Pragma("omp parallel for schedule(static)")                                              \
    for (int c = 0; c < C; ++c) {                                                             \
      int g = c / cpg;                                                                        \
      for (int b = 0; b < B; ++b) {                                                           \
        T *gim = grad_im + ((size_t)b * C + c) * H * W;                                       \
        const T *off_b = offset + ((size_t)b * dg + g) * 2 * kk * Ho * Wo;                    \
        const T *msk_b = mask ? mask + ((size_t)b * dg + g) * kk * Ho * Wo : NULL;            \
        for (int i = 0; i < kh; ++i)                                                          \
          for (int j = 0; j < kw; ++j) {                                                      \
            int k = i * kw + j;                                                               \
            const T *src = gcol + (((size_t)(c * kk + k) * B + b) * Ho) * Wo;                 \
            for (int ho = 0; ho < Ho; ++ho)                                                   \
              for (int wo = 0; wo < Wo; ++wo) {                                               \
                T h, w;                                                                       \
                sample_pos_##SUF(off_b, k, kk, Ho, Wo, ho, wo, i, j, sh, sw, ph, pw, dh, dw,  \
                                 scale_h, scale_w, &h, &w);                                   \
                T top = src[(size_t)ho * Wo + wo];                                            \
                if (msk_b) top *= msk_b[((size_t)k * Ho + ho) * Wo + wo];                     \
                int ch = (int)h, cw = (int)w; /* truncation, as the reference (:370-371) */   \
                for (int dy = -2; dy <= 2; ++dy)                                              \
                  for (int dx = -2; dx <= 2; ++dx) {                                          \
                    int yy = ch + dy, xx = cw + dx;                                           \
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W && fabs((double)(h - yy)) < 1 && \
                        fabs((double)(w - xx)) < 1) {                                         \
                      /* get_gradient_weight :117-143 */                                      \
                      T wt = 0;                                                               \
                      if (!(h <= -1 || h >= H || w <= -1 || w >= W)) {                        \
                        int hl = (int)floor((double)h), wl = (int)floor((double)w);           \
                        int hh_ = hl + 1, wh_ = wl + 1;                                       \
                        if (yy == hl && xx == wl) wt = (yy + 1 - h) * (xx + 1 - w);           \
                        if (yy == hl && xx == wh_) wt = (yy + 1 - h) * (w + 1 - xx);          \
                        if (yy == hh_ && xx == wl) wt = (h + 1 - yy) * (xx + 1 - w);          \
                        if (yy == hh_ && xx == wh_) wt = (h + 1 - yy) * (w + 1 - xx);         \
                      }                                                                       \
                      gim[(size_t)yy * W + xx] += wt * top;                                   \
                    }                                                                         \
                  }                                                                           \
              }                                                                               \
          }                                                                                   \
      }                                                                                       \
    }                                                                                         \
  }                                                                                           \
  /* col2im_coord: grad wrt offsets (and mask). …kernel.cu:486-549 / 551-615 / 972-1044.      \
     grad_offset (B, dg*2*kk, Ho, Wo) and grad_mask (B, dg*kk, Ho, Wo) are overwritten. */    \
  void dcn_col2im_coord_##SUF(const T *gcol, const T *x, const T *offset, const T *mask,      \
                              int B, int C, int H, int W, int Ho, int Wo, int kh, int kw,     \
                              int sh, int sw, int ph, int pw, int dh, int dw, float scale_h,  \
                              float scale_w, int dg, T *grad_offset, T *grad_mask) {          \
    int kk = kh * kw, cpg = C / dg;                                                           \
    for (int b = 0; b < B; ++b)                                                               \
      for (int g = 0; g < dg; ++g) {                                                          \
        const T *off_b = offset + ((size_t)b * dg + g) * 2 * kk * Ho * Wo;                    \
        const T *msk_b = mask ? mask + ((size_t)b * dg + g) * kk * Ho * Wo : NULL;            \
        for (int i = 0; i < kh; ++i)                                                          \
          for (int j = 0; j < kw; ++j) {                                                      \
            int k = i * kw + j;                                                               \
            _Pragma("omp parallel for schedule(static)")                                      \
            for (int ho = 0; ho < Ho; ++ho)                                                   \
              for (int wo = 0; wo < Wo; ++wo) {                                               \
                T h, w;                                                                       \
                sample_pos_##SUF(off_b, k, kk, Ho, Wo, ho, wo, i, j, sh, sw, ph, pw, dh, dw,  \
                                 scale_h, scale_w, &h, &w);                                   \
                int inside = !(h <= -1 || w <= -1 || h >= H || w >= W);                       \
                T m = msk_b ? msk_b[((size_t)k * Ho + ho) * Wo + wo] : (T)1;                  \
                T gh = 0, gw = 0, gm = 0;                                                     \
                for (int cc = 0; cc < cpg; ++cc) {                                            \
                  int c = g * cpg + cc;                                                       \
                  const T *im = x + ((size_t)b * C + c) * H * W;                              \
                  T gc = gcol[(((size_t)(c * kk + k) * B + b) * Ho + ho) * Wo + wo];          \
                  if (inside) {                                                               \
                    gm += gc * bilinear_##SUF(im, H, W, h, w);                                \
                    gh += coord_weight_##SUF(im, H, W, h, w, 0) * gc * m;                     \
                    gw += coord_weight_##SUF(im, H, W, h, w, 1) * gc * m;                     \
                  }                                                                           \
                }                                                                             \
                size_t o = (size_t)ho * Wo + wo;                                              \
                grad_offset[(((size_t)b * dg + g) * 2 * kk + 2 * k) * Ho * Wo + o] = gh;      \
                grad_offset[(((size_t)b * dg + g) * 2 * kk + 2 * k + 1) * Ho * Wo + o] = gw;  \
                if (grad_mask) grad_mask[(((size_t)b * dg + g) * kk + k) * Ho * Wo + o] = gm; \
              }                                                                               \
          }                                                                                   \
      }                                                                                       \
  }

DEFINE_DCN(float, f32)
DEFINE_DCN(double, f64)
