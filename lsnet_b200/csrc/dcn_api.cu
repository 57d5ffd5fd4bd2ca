// C-ABI entry points of the deformable convolutions as WHOLE operators -- one call per reference pybind function:
//   lsnet_dcn_forward          <- deform_conv_forward / modulated_deform_conv_forward / pyramid_deform_conv_forward
//   lsnet_dcn_backward_data    <- *_backward_input (+ the offset / mask gradients of modulated_deform_conv_backward)
//   lsnet_dcn_backward_weight  <- *_backward_parameters (+ the weight / bias half of modulated_deform_conv_backward)
// (mmdet/ops/dcn/src/deform_conv_ext.cpp:74-224, 227-250).  Each picks the fused tcgen05 kernel when the shape allows and
// the column-matrix path (gather -> GEMM, GEMM -> scatter) otherwise; the *_workspace_size queries tell the caller how
// many scratch bytes the call will need, so nothing is allocated inside.
#include <stdio.h>
#include <stdlib.h>

#include "../../include/lsnet_b200.h"
#include "common.cuh"
#include "dcn_common.cuh"
#include "lsnet_internal.h"

namespace lsn {
void dcn_fused_set(int on);
bool dcn_fused_supported(int C, int N, int kh, int kw, int dg, long long ldx, long long B, long long H, long long W,
                         long long Ho, long long Wo);
int dcn_fused_forward(const DcnGeom& g, const void* x, const float* offset, const float* mask, const void* Wp, int N,
                      const float* bias, int relu, void* out, long long ldc, int out_fp32, void* col, cudaStream_t st,
                      int grouped, double* gn_sums, int gn_G);
int gemm_bf16_gn(const void* A, long long lda, const void* Bw, long long ldb, void* out, long long ldc, int M, int N, int K,
                 const float* bias, int relu, int out_fp32, double* gn_sums, int gn_G, long long gn_hw, cudaStream_t st);
size_t dcn_fused_wgrad_partial_bytes(const DcnGeom& g, int M, long long ldw, int diag);
int dcn_fused_wgrad(const DcnGeom& g, const void* dy, long long ldy, int M, const void* x, const float* offset,
                    const float* mask, float* dW, long long ldw, float* partial, cudaStream_t st, int diag);

// Grouped weights (X-101-64x4d: groups = 64, C = 512 / 1024 / 2048) on the dedicated kernels: every 64-channel block holds
// whole groups (64 % (C / groups) == 0) and maps to the same 64 output channels (Cin == Cout).
static bool grouped_ok(const lsnet_dcn_desc* d, int N) {
  if (d->groups <= 1 || d->deformable_groups != 1 || d->kh * d->kw > 9 || d->C % 256 || N != d->C || d->C % d->groups) return false;
  const int cpg = d->C / d->groups;
  return cpg >= 1 && 64 % cpg == 0 && d->ldx % 8 == 0 && static_cast<long long>(d->B) * d->H * d->W < 2147483647LL &&
         static_cast<long long>(d->B) * d->Ho * d->Wo < 2147483647LL;
}

// LSNET_DETERMINISTIC=1 (or lsnet_set_deterministic): reductions that would otherwise use floating-point atomics in an
// order that varies from run to run take a fixed-order path (two-stage weight gradient, fp32 dX accumulation is unchanged)
static int g_deterministic = -1;
static bool deterministic() {
  if (g_deterministic < 0) {
    const char* e = getenv("LSNET_DETERMINISTIC");
    g_deterministic = (e && e[0] == '1') ? 1 : 0;
  }
  return g_deterministic != 0;
}

static bool wgrad_fused_ok(const lsnet_dcn_desc* d, int N) {
  return (d->C % 256) == 0 && N % 8 == 0 &&
         dcn_fused_supported(d->C, 16, d->kh, d->kw, d->deformable_groups, d->ldx, d->B, d->H, d->W, d->Ho, d->Wo);
}

static int check_desc(const char* who, const lsnet_dcn_desc* d) {
  if (!d) return set_error("%s: null descriptor", who);
  if (d->dtype != LSNET_DTYPE_BF16) return set_error("%s: dtype %d is not supported (LSNET_DTYPE_BF16 only)", who, d->dtype);
  if (d->groups < 1) return set_error("%s: groups = %d", who, d->groups);
  if (d->deformable_groups < 1 || d->C % d->deformable_groups || (d->C / d->deformable_groups) % 8 || d->ldx % 8)
    return set_error("%s: need C / deformable_groups %% 8 == 0 and a 16-byte aligned pixel pitch (C=%d dg=%d ldx=%lld)",
                     who, d->C, d->deformable_groups, d->ldx);
  if (d->kh < 1 || d->kw < 1 || d->Ho < 0 || d->Wo < 0 || d->B < 0) return set_error("%s: bad extents", who);
  return 0;
}

static DcnGeom geom_of(const lsnet_dcn_desc* d, long long ldo, long long ldm, long long ldcol, const float* mask) {
  return DcnGeom{d->B, d->H, d->W, d->C, d->Ho, d->Wo, d->kh, d->kw, d->stride_h, d->stride_w, d->pad_h, d->pad_w,
                 d->dil_h, d->dil_w, d->scale_h, d->scale_w, d->deformable_groups, d->ldx, ldo, ldm, ldcol,
                 (mask && d->mask_logits) ? 1 : 0};
}

static size_t col_bytes(const lsnet_dcn_desc* d) {
  return static_cast<size_t>(d->B) * d->Ho * d->Wo * d->kh * d->kw * d->C * 2;
}
}  // namespace lsn

using namespace lsn;

extern "C" void lsnet_dcn_fused_enable(int on) { dcn_fused_set(on); }
extern "C" void lsnet_set_deterministic(int on) { g_deterministic = on ? 1 : 0; }

extern "C" int lsnet_dcn_grouped_supported(const lsnet_dcn_desc* d, int N) { return d && grouped_ok(d, N) ? 1 : 0; }

static int need_grouped(const char* who, const lsnet_dcn_desc* d, int N) {
  if (d->groups > 1 && !grouped_ok(d, N))
    return set_error("%s: groups = %d with C = %d, N = %d is outside the grouped kernels (need C == N, C %% 256 == 0, "
                     "64 %% (C / groups) == 0, deformable_groups == 1): expand the weight to a dense block-diagonal pack",
                     who, d->groups, d->C, N);
  return 0;
}

extern "C" size_t lsnet_dcn_forward_workspace_size(const lsnet_dcn_desc* d, int N) {
  if (!d) return 0;
  if (d->groups > 1) return 0;
  if (dcn_fused_supported(d->C, N, d->kh, d->kw, d->deformable_groups, d->ldx, d->B, d->H, d->W, d->Ho, d->Wo)) return 0;
  return col_bytes(d);
}

extern "C" int lsnet_dcn_forward(const lsnet_dcn_desc* d, const void* x, const float* offset, long long ldo,
                                 const float* mask, long long ldm, const void* Wp, int N, const float* bias, int relu,
                                 void* out, long long ldc, int out_fp32, void* col_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  return lsnet_dcn_forward_gn(d, x, offset, ldo, mask, ldm, Wp, N, bias, relu, out, ldc, out_fp32, col_out, workspace,
                              workspace_bytes, nullptr, 0, stream);
}

extern "C" int lsnet_dcn_forward_gn(const lsnet_dcn_desc* d, const void* x, const float* offset, long long ldo,
                                    const float* mask, long long ldm, const void* Wp, int N, const float* bias, int relu,
                                    void* out, long long ldc, int out_fp32, void* col_out, void* workspace,
                                    size_t workspace_bytes, double* gn_sums, int gn_groups, void* stream) {
  if (int rc = check_desc("lsnet_dcn_forward", d)) return rc;
  if (gn_sums && (d->groups > 1 || gn_groups < 1 || N % gn_groups || (N / gn_groups) % 8))
    return set_error("lsnet_dcn_forward_gn: GroupNorm statistics need groups == 1, N %% G == 0, (N / G) %% 8 == 0 (N=%d G=%d)", N, gn_groups);
  if (d->B == 0 || d->Ho == 0 || d->Wo == 0) return 0;
  if (N <= 0 || N % 16) return set_error("lsnet_dcn_forward: N (= rows of the packed weight) must be a positive multiple of 16, got %d", N);
  if (ldc % (out_fp32 ? 4 : 8)) return set_error("lsnet_dcn_forward: output pitch must be 16-byte aligned");
  if (int rc = need_grouped("lsnet_dcn_forward", d, N)) return rc;
  const long long K = static_cast<long long>(d->kh) * d->kw * d->C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->groups > 1) {
    if (col_out) return set_error("lsnet_dcn_forward: the grouped kernel has no column side output");
    const DcnGeom g = geom_of(d, ldo, ldm, K, mask);
    return dcn_fused_forward(g, x, offset, mask, Wp, N, bias, relu, out, ldc, out_fp32, nullptr, st, 1, nullptr, 0);
  }
  if (dcn_fused_supported(d->C, N, d->kh, d->kw, d->deformable_groups, d->ldx, d->B, d->H, d->W, d->Ho, d->Wo)) {
    const DcnGeom g = geom_of(d, ldo, ldm, K, mask);
    return dcn_fused_forward(g, x, offset, mask, Wp, N, bias, relu, out, ldc, out_fp32, col_out, st, 0, gn_sums, gn_groups);
  }
  void* col = col_out ? col_out : workspace;
  if (!col || (!col_out && workspace_bytes < col_bytes(d)))
    return set_error("lsnet_dcn_forward: this shape takes the column-matrix path and needs %zu workspace bytes (got %zu)",
                     col_bytes(d), workspace_bytes);
  if (int rc = lsnet_dcn_im2col_bf16(x, d->B, d->H, d->W, d->C, d->ldx, offset, ldo, mask, ldm, d->Ho, d->Wo, d->kh,
                                     d->kw, d->stride_h, d->stride_w, d->pad_h, d->pad_w, d->dil_h, d->dil_w,
                                     d->scale_h, d->scale_w, d->deformable_groups, col, K, d->mask_logits, stream))
    return rc;
  return gemm_bf16_gn(col, K, Wp, K, out, ldc, d->B * d->Ho * d->Wo, N, static_cast<int>(K), bias, relu, out_fp32, gn_sums,
                      gn_groups, static_cast<long long>(d->Ho) * d->Wo, st);
}

extern "C" size_t lsnet_dcn_backward_data_workspace_size(const lsnet_dcn_desc* d, int N) {
  (void)N;
  return d ? col_bytes(d) : 0;
}

extern "C" int lsnet_dcn_backward_data(const lsnet_dcn_desc* d, const void* dy, long long ldy, int N, const void* Wt,
                                       const void* x, const float* offset, long long ldo, const float* mask,
                                       long long ldm, void* dx, long long lddx, int dx_fp32, float* doffset,
                                       long long lddo, float* dmask, long long lddm, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (int rc = check_desc("lsnet_dcn_backward_data", d)) return rc;
  if (d->B == 0 || d->Ho == 0 || d->Wo == 0) return 0;
  if (N <= 0 || N % 8 || ldy % 8) return set_error("lsnet_dcn_backward_data: N and the dY pitch must be multiples of 8");
  if (int rc = need_grouped("lsnet_dcn_backward_data", d, N)) return rc;
  const long long K = static_cast<long long>(d->kh) * d->kw * d->C;
  if (!workspace || workspace_bytes < col_bytes(d))
    return set_error("lsnet_dcn_backward_data: needs %zu workspace bytes (got %zu)", col_bytes(d), workspace_bytes);
  if (d->groups > 1) {
    // dCol[p, tap*C + blk*64 + j] = sum_{o in block blk} dY[p, blk*64 + o] * Wt[tap*C + blk*64 + j, o]: one 64 x 64 product
    // per (tap, block) instead of the dense GEMM's groups-times redundant one
    if (int rc = lsnet_gemm_blockdiag_bf16(dy, ldy, Wt, workspace, K, d->B * d->Ho * d->Wo, static_cast<int>(K), d->C / 64,
                                           stream))
      return rc;
  } else if (int rc = lsnet_gemm_bf16(dy, ldy, Wt, N, workspace, K, d->B * d->Ho * d->Wo, static_cast<int>(K), N, nullptr,
                                      0, 0, stream)) {
    // dCol[p, tap*C + c] = sum_n dY[p, n] * Wt[tap*C + c, n]
    return rc;
  }
  return lsnet_dcn_col2im_bf16(workspace, K, x, d->B, d->H, d->W, d->C, d->ldx, offset, ldo, mask, ldm, d->Ho, d->Wo,
                               d->kh, d->kw, d->stride_h, d->stride_w, d->pad_h, d->pad_w, d->dil_h, d->dil_w,
                               d->scale_h, d->scale_w, d->deformable_groups, dx, lddx, dx_fp32, doffset, lddo, dmask,
                               lddm, d->mask_logits, stream);
}

extern "C" size_t lsnet_dcn_backward_weight_workspace_size(const lsnet_dcn_desc* d, int N, int have_col) {
  if (!d) return 0;
  if (d->groups > 1) {
    if (!deterministic() || !grouped_ok(d, N)) return 0;
    const long long K = static_cast<long long>(d->kh) * d->kw * d->C;
    return dcn_fused_wgrad_partial_bytes(geom_of(d, 0, 0, K, nullptr), N, static_cast<long long>(d->kh) * d->kw * 256, 1);
  }
  if (have_col) return 0;
  if (wgrad_fused_ok(d, N)) {
    if (!deterministic()) return 0;
    const long long K = static_cast<long long>(d->kh) * d->kw * d->C;
    return dcn_fused_wgrad_partial_bytes(geom_of(d, 0, 0, K, nullptr), N, K, 0);
  }
  return col_bytes(d);
}

extern "C" int lsnet_dcn_backward_weight(const lsnet_dcn_desc* d, const void* dy, long long ldy, int N, const void* x,
                                         const float* offset, long long ldo, const float* mask, long long ldm,
                                         const void* col_saved, float* dW, long long lddw, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  if (int rc = check_desc("lsnet_dcn_backward_weight", d)) return rc;
  if (d->B == 0 || d->Ho == 0 || d->Wo == 0) return 0;
  const long long K = static_cast<long long>(d->kh) * d->kw * d->C;
  if (N <= 0 || N % 8 || ldy % 8 || lddw % 4) return set_error("lsnet_dcn_backward_weight: N, dY pitch %% 8 and dW pitch %% 4 required");
  if (int rc = need_grouped("lsnet_dcn_backward_weight", d, N)) return rc;
  if (d->groups > 1) {
    // grouped: dW is the compact block-diagonal form [C, kh*kw, 256] (row o: its 256-channel tile of every tap); only the
    // tiles on the diagonal are computed, the columns are re-sampled in the kernel
    if (lddw != static_cast<long long>(d->kh) * d->kw * 256)
      return set_error("lsnet_dcn_backward_weight: grouped dW must be [C, kh*kw*256] (pitch %lld)", lddw);
    const DcnGeom g = geom_of(d, ldo, ldm, K, mask);
    float* partial = nullptr;
    if (deterministic()) {
      const size_t need = dcn_fused_wgrad_partial_bytes(g, N, lddw, 1);
      if (!workspace || workspace_bytes < need)
        return set_error("lsnet_dcn_backward_weight: deterministic mode needs %zu workspace bytes (got %zu)", need, workspace_bytes);
      partial = static_cast<float*>(workspace);
    }
    return dcn_fused_wgrad(g, dy, ldy, N, x, offset, mask, dW, lddw, partial, static_cast<cudaStream_t>(stream), 1);
  }
  const void* col = col_saved;
  if (!col && wgrad_fused_ok(d, N)) {
    // columns re-sampled inside the tcgen05 weight-gradient kernel: nothing of the column matrix touches HBM
    const DcnGeom g = geom_of(d, ldo, ldm, K, mask);
    float* partial = nullptr;
    if (deterministic()) {
      const size_t need = dcn_fused_wgrad_partial_bytes(g, N, lddw, 0);
      if (!workspace || workspace_bytes < need)
        return set_error("lsnet_dcn_backward_weight: deterministic mode needs %zu workspace bytes (got %zu)", need, workspace_bytes);
      partial = static_cast<float*>(workspace);
    }
    return dcn_fused_wgrad(g, dy, ldy, N, x, offset, mask, dW, lddw, partial, static_cast<cudaStream_t>(stream), 0);
  }
  if (!col) {
    if (!workspace || workspace_bytes < col_bytes(d))
      return set_error("lsnet_dcn_backward_weight: without saved columns the call needs %zu workspace bytes (got %zu)",
                       col_bytes(d), workspace_bytes);
    if (int rc = lsnet_dcn_im2col_bf16(x, d->B, d->H, d->W, d->C, d->ldx, offset, ldo, mask, ldm, d->Ho, d->Wo, d->kh,
                                       d->kw, d->stride_h, d->stride_w, d->pad_h, d->pad_w, d->dil_h, d->dil_w,
                                       d->scale_h, d->scale_w, d->deformable_groups, workspace, K, d->mask_logits,
                                       stream))
      return rc;
    col = workspace;
  }
  return lsnet_gemm_tn_bf16(dy, ldy, col, K, dW, lddw, d->B * d->Ho * d->Wo, N, static_cast<int>(K), stream);
}
