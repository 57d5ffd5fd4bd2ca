// Loss kernels (sm_100a): cross-IOU (bbox / polygon / keypoint) forward + backward, directional regression targets,
// sigmoid focal loss forward + backward.  Streaming / latency-bound: algorithmic bytes N*(4D+4) for the fused
// cross-IOU (pred fp32 + assigned-GT index; anchors come from the row index, targets from the compact GT table),
// N*(4C+8) for focal.  Row arithmetic lives in loss_math.cuh (shared with the host parity harness).
#include "common.cuh"
#include "loss_math.cuh"
#include "lsnet_internal.h"

namespace lsn {

// ------------------------------------------------------------------------------------------------------------
// Dense form: the drop-in for CrossIOULoss.forward(pred, target, weight, anchor_pts=, bbox_gt=, pos_inds=, vs=)
// (cross_iou_loss.py:146-172).  weight_row = weight.mean(-1).  row_loss[n] = weight_row[n] * loss_n.
// ------------------------------------------------------------------------------------------------------------
__global__ void cross_iou_dense_fwd_kernel(int type, const float* __restrict__ pred, const float* __restrict__ target,
                                           const uint8_t* __restrict__ sel, const float* __restrict__ wrow,
                                           const float* __restrict__ anchor, const float* __restrict__ bbox_gt,
                                           const float* __restrict__ vs, int N, int D, int L, float eps, float alpha,
                                           int pstride, float* __restrict__ row_loss) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float w = wrow ? wrow[n] : 1.f;
  float out = 0.f;
  if (w > 0.f) {
    float p[kMaxD], t[kMaxD];
    uint8_t s[kMaxD];
    for (int d = 0; d < D; ++d) {
      p[d] = pred[static_cast<long long>(n) * D + d];
      t[d] = target[static_cast<long long>(n) * D + d];
      s[d] = sel[static_cast<long long>(n) * D + d];
    }
    float a[2] = {0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
    if (anchor) { a[0] = anchor[2 * n]; a[1] = anchor[2 * n + 1]; }
    if (bbox_gt) for (int e = 0; e < 4; ++e) g[e] = bbox_gt[4 * n + e];
    out = w * cross_iou_row(type, p, t, s, D, a, g, vs ? vs + static_cast<long long>(n) * L : nullptr, eps, alpha,
                            pstride, nullptr);
  }
  row_loss[n] = out;
}

__global__ void cross_iou_dense_bwd_kernel(int type, const float* __restrict__ pred, const float* __restrict__ target,
                                           const uint8_t* __restrict__ sel, const float* __restrict__ wrow,
                                           const float* __restrict__ anchor, const float* __restrict__ bbox_gt,
                                           const float* __restrict__ vs, int N, int D, int L, float eps, float alpha,
                                           int pstride, const float* __restrict__ scale_ptr,
                                           float* __restrict__ dpred) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float w = wrow ? wrow[n] : 1.f;
  if (!(w > 0.f)) return;   // dpred was zero-filled
  float p[kMaxD], t[kMaxD], gr[kMaxD];
  uint8_t s[kMaxD];
  for (int d = 0; d < D; ++d) {
    p[d] = pred[static_cast<long long>(n) * D + d];
    t[d] = target[static_cast<long long>(n) * D + d];
    s[d] = sel[static_cast<long long>(n) * D + d];
  }
  float a[2] = {0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
  if (anchor) { a[0] = anchor[2 * n]; a[1] = anchor[2 * n + 1]; }
  if (bbox_gt) for (int e = 0; e < 4; ++e) g[e] = bbox_gt[4 * n + e];
  cross_iou_row(type, p, t, s, D, a, g, vs ? vs + static_cast<long long>(n) * L : nullptr, eps, alpha, pstride, gr);
  const float sc = w * (*scale_ptr);
  for (int d = 0; d < D; ++d) dpred[static_cast<long long>(n) * D + d] = sc * gr[d];
}

// ------------------------------------------------------------------------------------------------------------
// Fused form used by LSHead.loss: per level, rows = (b, h, w) of an NHWC prediction map; the target row is
// rebuilt on the fly from the point's assigned GT (lsnet_head.py:1064-1102: get_bbox_gt_reg + division by
// normalize_term = point_base_scale * stride, all exact power-of-two scalings).
// ------------------------------------------------------------------------------------------------------------
struct FusedArgs {
  int type, D, NP, L;              // D = 4*NP
  int B, Hl, Wl, level_off, Gmax;
  long long ldp, assign_ld;
  float stride, base_scale, eps, alpha;
  int pstride;
};

template <bool BWD>
__global__ void cross_iou_fused_kernel(const FusedArgs a, const float* __restrict__ pred,
                                       const int* __restrict__ assign, const float* __restrict__ gt_pts,
                                       const float* __restrict__ gt_bbox, const float* __restrict__ gt_vs,
                                       float* __restrict__ row_loss, const float* __restrict__ scale_ptr,
                                       float* __restrict__ dpred) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int Pl = a.Hl * a.Wl;
  if (n >= a.B * Pl) return;
  const int b = n / Pl, pix = n % Pl;
  const int gi = assign[static_cast<long long>(b) * a.assign_ld + a.level_off + pix];
  if (gi < 0) {
    if (!BWD) row_loss[n] = 0.f;
    return;
  }
  const float nt = a.base_scale * a.stride;
  const float anchor_raw[2] = {static_cast<float>(pix % a.Wl) * a.stride, static_cast<float>(pix / a.Wl) * a.stride};
  const float* gp = gt_pts + (static_cast<long long>(b) * a.Gmax + gi) * 2 * a.NP;
  float p[kMaxD], t[kMaxD], gr[kMaxD];
  uint8_t s[kMaxD];
  directional_target_row(gp, a.NP, anchor_raw, true, t, s);
  for (int d = 0; d < a.D; ++d) {
    p[d] = (pred[static_cast<long long>(n) * a.ldp + d] * a.stride) / nt;
    t[d] = t[d] / nt;
  }
  const float anchor[2] = {anchor_raw[0] / nt, anchor_raw[1] / nt};
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (gt_bbox) for (int e = 0; e < 4; ++e) g[e] = gt_bbox[(static_cast<long long>(b) * a.Gmax + gi) * 4 + e] / nt;
  const float* vs = gt_vs ? gt_vs + (static_cast<long long>(b) * a.Gmax + gi) * a.L : nullptr;
  if (!BWD) {
    row_loss[n] = cross_iou_row(a.type, p, t, s, a.D, anchor, g, vs, a.eps, a.alpha, a.pstride, nullptr);
  } else {
    cross_iou_row(a.type, p, t, s, a.D, anchor, g, vs, a.eps, a.alpha, a.pstride, gr);
    const float sc = (*scale_ptr) * (a.stride / nt);
    for (int d = 0; d < a.D; ++d) dpred[static_cast<long long>(n) * a.ldp + d] = sc * gr[d];
  }
}

// Directional targets, dense (LSHead.get_bbox_gt_reg / get_poly_gt_reg, lsnet_head.py:402-454).
__global__ void directional_targets_kernel(const float* __restrict__ gt_rows, const float* __restrict__ anchor,
                                           const float* __restrict__ wrow, int N, int NP, float* __restrict__ target,
                                           uint8_t* __restrict__ sel) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t[kMaxD];
  uint8_t s[kMaxD];
  const float a[2] = {anchor[2 * n], anchor[2 * n + 1]};
  float gp[kMaxD / 2];
  for (int j = 0; j < 2 * NP; ++j) gp[j] = gt_rows[static_cast<long long>(n) * 2 * NP + j];
  directional_target_row(gp, NP, a, wrow[n] > 0.f, t, s);
  for (int d = 0; d < 4 * NP; ++d) {
    target[static_cast<long long>(n) * 4 * NP + d] = t[d];
    sel[static_cast<long long>(n) * 4 * NP + d] = s[d];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Sigmoid focal loss (sigmoid_focal_loss_cuda.cu:23-97).  labels int32 (num_classes = background), per-row weight.
// Forward writes per-block partial sums of weight*loss (deterministic two-stage reduction); backward is
// element-wise: dlogit = scale * weight[n] * dloss.
// ------------------------------------------------------------------------------------------------------------
constexpr int FOCAL_THREADS = 256;
__global__ void __launch_bounds__(FOCAL_THREADS)
focal_fwd_kernel(const float* __restrict__ logits, long long ldl, const int* __restrict__ labels,
                 const float* __restrict__ weight, long long N, int C, float gamma, float alpha,
                 float* __restrict__ partial) {
  float acc = 0.f;
  const long long total = N * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / C;
    const int d = static_cast<int>(i % C);
    const float w = weight ? weight[n] : 1.f;
    acc += w * focal_elem(logits[n * ldl + d], labels[n], d, gamma, alpha, nullptr);
  }
  __shared__ float red[FOCAL_THREADS / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < FOCAL_THREADS / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(FOCAL_THREADS)
focal_bwd_kernel(const float* __restrict__ logits, long long ldl, const int* __restrict__ labels,
                 const float* __restrict__ weight, long long N, int C, float gamma, float alpha,
                 const float* __restrict__ scale_ptr, float* __restrict__ dlogits, long long ldd) {
  const long long total = N * C;
  const float sc = *scale_ptr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / C;
    const int d = static_cast<int>(i % C);
    const float w = weight ? weight[n] : 1.f;
    float g;
    focal_elem(logits[n * ldl + d], labels[n], d, gamma, alpha, &g);
    dlogits[n * ldd + d] = sc * w * g;
  }
}

}  // namespace lsn

using namespace lsn;

static inline int nblk(long long n, int t) { return static_cast<int>((n + t - 1) / t); }

extern "C" int lsnet_cross_iou_fwd(int loss_type, const float* pred, const float* target, const unsigned char* pos_inds,
                                   const float* weight_row, const float* anchor_pts, const float* bbox_gt,
                                   const float* vs, int N, int D, int L, float eps, float alpha, int stride,
                                   float* row_loss, void* stream) {
  if (N <= 0) return 0;
  if (D > kMaxD || D % 4) return set_error("lsnet_cross_iou_fwd: D=%d unsupported (multiple of 4, <= %d)", D, kMaxD);
  cross_iou_dense_fwd_kernel<<<nblk(N, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      loss_type, pred, target, pos_inds, weight_row, anchor_pts, bbox_gt, vs, N, D, L, eps, alpha, stride, row_loss);
  return check_launch("cross_iou_dense_fwd");
}

extern "C" int lsnet_cross_iou_bwd(int loss_type, const float* pred, const float* target, const unsigned char* pos_inds,
                                   const float* weight_row, const float* anchor_pts, const float* bbox_gt,
                                   const float* vs, int N, int D, int L, float eps, float alpha, int stride,
                                   const float* scale, float* dpred, void* stream) {
  if (N <= 0) return 0;
  if (D > kMaxD || D % 4) return set_error("lsnet_cross_iou_bwd: D=%d unsupported", D);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(dpred, 0, sizeof(float) * static_cast<size_t>(N) * D, st);
  cross_iou_dense_bwd_kernel<<<nblk(N, 128), 128, 0, st>>>(loss_type, pred, target, pos_inds, weight_row, anchor_pts,
                                                           bbox_gt, vs, N, D, L, eps, alpha, stride, scale, dpred);
  return check_launch("cross_iou_dense_bwd");
}

extern "C" int lsnet_cross_iou_level(int loss_type, int backward, const float* pred, long long ldp, int D,
                                     const int* assign, long long assign_ld, int level_off, int B, int Hl, int Wl,
                                     float stride, float base_scale, const float* gt_pts, const float* gt_bbox,
                                     const float* gt_vs, int Gmax, int NP, int L, float eps, float alpha,
                                     int pstride, float* row_loss, const float* scale, float* dpred, void* stream) {
  const long long N = static_cast<long long>(B) * Hl * Wl;
  if (N <= 0) return 0;
  if (D != 4 * NP || D > kMaxD) return set_error("lsnet_cross_iou_level: D=%d NP=%d inconsistent", D, NP);
  FusedArgs a{loss_type, D, NP, L, B, Hl, Wl, level_off, Gmax, ldp, assign_ld, stride, base_scale, eps, alpha, pstride};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!backward) {
    cross_iou_fused_kernel<false><<<nblk(N, 128), 128, 0, st>>>(a, pred, assign, gt_pts, gt_bbox, gt_vs, row_loss,
                                                               nullptr, nullptr);
  } else {
    cudaMemsetAsync(dpred, 0, sizeof(float) * static_cast<size_t>(N) * ldp, st);
    cross_iou_fused_kernel<true><<<nblk(N, 128), 128, 0, st>>>(a, pred, assign, gt_pts, gt_bbox, gt_vs, nullptr,
                                                              scale, dpred);
  }
  return check_launch("cross_iou_fused");
}

extern "C" int lsnet_directional_targets(const float* gt_rows, const float* anchor_pts, const float* weight_row, int N,
                                         int NP, float* target, unsigned char* pos_inds, void* stream) {
  if (N <= 0) return 0;
  if (4 * NP > kMaxD) return set_error("lsnet_directional_targets: NP=%d too large", NP);
  directional_targets_kernel<<<nblk(N, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(gt_rows, anchor_pts,
                                                                                         weight_row, N, NP, target,
                                                                                         pos_inds);
  return check_launch("directional_targets");
}

extern "C" int lsnet_focal_partial_count(long long N, int C) {
  long long blocks = (N * C + FOCAL_THREADS - 1) / FOCAL_THREADS;
  return static_cast<int>(blocks < 1184 ? (blocks < 1 ? 1 : blocks) : 1184);   // 148 SMs x 8 resident CTAs
}

extern "C" int lsnet_focal_fwd(const float* logits, long long ldl, const int* labels, const float* weight,
                               long long N, int C, float gamma, float alpha, float* partial, void* stream) {
  if (N <= 0) return 0;
  focal_fwd_kernel<<<lsnet_focal_partial_count(N, C), FOCAL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ldl, labels, weight, N, C, gamma, alpha, partial);
  return check_launch("focal_fwd");
}

extern "C" int lsnet_focal_bwd(const float* logits, long long ldl, const int* labels, const float* weight,
                               long long N, int C, float gamma, float alpha, const float* scale, float* dlogits,
                               long long ldd, void* stream) {
  if (N <= 0) return 0;
  focal_bwd_kernel<<<lsnet_focal_partial_count(N, C), FOCAL_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, ldl, labels, weight, N, C, gamma, alpha, scale, dlogits, ldd);
  return check_launch("focal_bwd");
}

// ---- host-side parity hooks: run the exact row arithmetic on the CPU (no GPU needed) -------------------------
extern "C" float lsnet_host_cross_iou_row(int loss_type, const float* pred, const float* target,
                                          const unsigned char* pos_inds, int D, const float* anchor,
                                          const float* bbox_gt, const float* vs, float eps, float alpha, int stride,
                                          float* grad) {
  return cross_iou_row(loss_type, pred, target, pos_inds, D, anchor, bbox_gt, vs, eps, alpha, stride, grad);
}
extern "C" float lsnet_host_focal_elem(float x, int t, int d, float gamma, float alpha, float* grad) {
  return focal_elem(x, t, d, gamma, alpha, grad);
}
extern "C" void lsnet_host_directional_target_row(const float* gt, int NP, const float* anchor, int positive,
                                                  float* target, unsigned char* sel) {
  directional_target_row(gt, NP, anchor, positive != 0, target, sel);
}
