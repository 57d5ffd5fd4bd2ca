// Landmark-target assignment on the GPU, batched over images (sm_100a).  Integer outputs are bit-exact with the
// reference on tie-free inputs; all distance / IoU arithmetic is plain fp32 in PyTorch's op order (explicit
// __fadd_rn/__fmul_rn/__fdiv_rn so nvcc never contracts to FMA).
//
//   centroid_*   CentroidAssigner.assign   (mmdet/core/bbox/assigners/centroid_assigner.py:26-93), iou_type='center'
//   atss_*       ATSSAssigner.assign       (mmdet/core/bbox/assigners/atss_assigner.py:29-164) + bbox_overlaps
//                                          (iou_calculators/iou2d_calculator.py:82-130)
//   targets      LSHead._target_single label / weight scatter (lsnet_head.py:834-898) and the per-image positive
//                counts behind num_total_pos (lsnet_head.py:984)
//   pred_boxes   extreme_points2bbox / vectors2bbox + centre shift (lsnet_head.py:321-370, 1333-1361)
//
// Points are never materialised: point (level l, h, w) = (w*stride_l, h*stride_l, stride_l) as PointGenerator
// builds them (core/anchor/point_generator.py:17-25); validity = (h < valid_h[b][l] && w < valid_w[b][l]) from
// the image's pad_shape (lsnet_head.py:781-792).
#include "common.cuh"
#include <stdint.h>
#include "lsnet_internal.h"

namespace lsn {

constexpr int kMaxLevels = 8;
struct Levels {
  int n, total;
  int H[kMaxLevels], W[kMaxLevels], off[kMaxLevels];
  float stride[kMaxLevels];
};

constexpr int ASSIGN_THREADS = 256;
constexpr float kInf = 1e8f;

struct MinPair {
  float d;
  int i;
};
__device__ __forceinline__ bool lex_less(float d0, int i0, float d1, int i1) { return d0 < d1 || (d0 == d1 && i0 < i1); }
__device__ __forceinline__ MinPair warp_min(MinPair v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float d = __shfl_xor_sync(0xffffffffu, v.d, o);
    const int i = __shfl_xor_sync(0xffffffffu, v.i, o);
    if (lex_less(d, i, v.d, v.i)) { v.d = d; v.i = i; }
  }
  return v;
}
__device__ MinPair block_min(MinPair v, MinPair* smem) {
  v = warp_min(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  MinPair r = smem[0];
  for (int k = 1; k < ASSIGN_THREADS / 32; ++k)
    if (lex_less(smem[k].d, smem[k].i, r.d, r.i)) r = smem[k];
  return r;
}

// ------------------------------------------------------------------------------------------------ centroid
// grid (Gmax, B): nearest valid grid point of the GT's scale level.
__global__ void __launch_bounds__(ASSIGN_THREADS)
centroid_nearest_kernel(const Levels lv, const int* __restrict__ valid_hw, const float* __restrict__ gt_bbox,
                        const int* __restrict__ gt_count, int Gmax, float scale, float* __restrict__ best_d,
                        int* __restrict__ best_i) {
  __shared__ MinPair red[ASSIGN_THREADS / 32];
  const int g = blockIdx.x, b = blockIdx.y;
  if (g >= gt_count[b]) return;
  const float* bb = gt_bbox + (static_cast<long long>(b) * Gmax + g) * 4;
  const float cx = __fdiv_rn(__fadd_rn(bb[0], bb[2]), 2.f), cy = __fdiv_rn(__fadd_rn(bb[1], bb[3]), 2.f);
  const float gw = fmaxf(__fadd_rn(bb[2], -bb[0]), 1e-6f), gh = fmaxf(__fadd_rn(bb[3], -bb[1]), 1e-6f);
  // level of the GT: trunc((log2(w/scale) + log2(h/scale)) / 2), clamped to the levels present (:61-65)
  const int lvl_min = static_cast<int>(log2f(lv.stride[0])), lvl_max = static_cast<int>(log2f(lv.stride[lv.n - 1]));
  int gl = static_cast<int>(__fdiv_rn(__fadd_rn(log2f(__fdiv_rn(gw, scale)), log2f(__fdiv_rn(gh, scale))), 2.f));
  gl = min(max(gl, lvl_min), lvl_max);
  MinPair best{kInf, 0x7fffffff};
  for (int l = 0; l < lv.n; ++l) {
    if (static_cast<int>(log2f(lv.stride[l])) != gl) continue;
    const int vh = valid_hw[(b * lv.n + l) * 2], vw = valid_hw[(b * lv.n + l) * 2 + 1];
    const int P = lv.H[l] * lv.W[l];
    for (int p = threadIdx.x; p < P; p += ASSIGN_THREADS) {
      const int h = p / lv.W[l], w = p % lv.W[l];
      if (h >= vh || w >= vw) continue;
      const float dx = __fdiv_rn(__fadd_rn(static_cast<float>(w) * lv.stride[l], -cx), gw);
      const float dy = __fdiv_rn(__fadd_rn(static_cast<float>(h) * lv.stride[l], -cy), gh);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (lex_less(d, lv.off[l] + p, best.d, best.i)) { best.d = d; best.i = lv.off[l] + p; }
    }
  }
  best = block_min(best, red);
  if (threadIdx.x == 0) {
    best_d[b * Gmax + g] = best.d;
    best_i[b * Gmax + g] = best.i;
  }
}

// grid (B): a point claimed by several GTs keeps the nearest (first GT on ties) (:76-80).  assign pre-filled -1.
__global__ void centroid_resolve_kernel(const float* __restrict__ best_d, const int* __restrict__ best_i,
                                        const int* __restrict__ gt_count, int Gmax, int* __restrict__ assign,
                                        long long assign_ld) {
  const int b = blockIdx.x;
  const int G = gt_count[b];
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float d = best_d[b * Gmax + g];
    const int i = best_i[b * Gmax + g];
    if (!(d < kInf) || i == 0x7fffffff) continue;
    bool win = true;
    for (int o = 0; o < G; ++o) {
      if (o == g || best_i[b * Gmax + o] != i) continue;
      const float od = best_d[b * Gmax + o];
      if (od < d || (od == d && o < g)) { win = false; break; }
    }
    if (win) assign[static_cast<long long>(b) * assign_ld + i] = g;
  }
}

// ------------------------------------------------------------------------------------------------ ATSS
__device__ __forceinline__ float iou_fp32(const float* a, const float* g) {
  const float ltx = fmaxf(a[0], g[0]), lty = fmaxf(a[1], g[1]);
  const float rbx = fminf(a[2], g[2]), rby = fminf(a[3], g[3]);
  const float w = fmaxf(__fadd_rn(rbx, -ltx), 0.f), h = fmaxf(__fadd_rn(rby, -lty), 0.f);
  const float overlap = __fmul_rn(w, h);
  const float area1 = __fmul_rn(__fadd_rn(a[2], -a[0]), __fadd_rn(a[3], -a[1]));
  const float area2 = __fmul_rn(__fadd_rn(g[2], -g[0]), __fadd_rn(g[3], -g[1]));
  const float uni = fmaxf(__fadd_rn(__fadd_rn(area1, area2), -overlap), 1e-6f);
  return __fdiv_rn(overlap, uni);
}

constexpr int kMaxCand = kMaxLevels * 16;

// grid (Gmax, B).  boxes: [B, total, 4] predicted init boxes.  Positive candidates are merged per point with a
// 64-bit atomicMax on (IoU bits, ~gt index): highest IoU wins, first GT on ties (:144-152).
__global__ void __launch_bounds__(ASSIGN_THREADS)
atss_candidates_kernel(const Levels lv, const int* __restrict__ valid_hw, const float* __restrict__ boxes,
                       const float* __restrict__ gt_bbox, const int* __restrict__ gt_count, int Gmax, int topk,
                       unsigned long long* __restrict__ best_key, int dist_cap) {
  extern __shared__ float dist_s[];      // [dist_cap] centre distances of the level being processed
  __shared__ MinPair red[ASSIGN_THREADS / 32];
  __shared__ int cand[kMaxCand];
  __shared__ float cand_iou[kMaxCand];
  __shared__ int ncand;
  __shared__ float thr_s;
  const int g = blockIdx.x, b = blockIdx.y;
  if (g >= gt_count[b]) return;
  float gb[4];
  for (int e = 0; e < 4; ++e) gb[e] = gt_bbox[(static_cast<long long>(b) * Gmax + g) * 4 + e];
  const float gcx = __fdiv_rn(__fadd_rn(gb[0], gb[2]), 2.f), gcy = __fdiv_rn(__fadd_rn(gb[1], gb[3]), 2.f);
  const float* bx = boxes + static_cast<long long>(b) * lv.total * 4;
  if (threadIdx.x == 0) ncand = 0;
  __syncthreads();
  for (int l = 0; l < lv.n; ++l) {
    const int vh = valid_hw[(b * lv.n + l) * 2], vw = valid_hw[(b * lv.n + l) * 2 + 1];
    const int P = lv.H[l] * lv.W[l];
    const int k_l = min(topk, vh * vw);
    float pd = -1.f;     // previously selected (distance, index): next pick is the lexicographic successor
    int pi = -1;
    const bool staged = P <= dist_cap;
    if (staged) {
      // centre distances of this level once into shared memory (invalid points: +inf); the k_l selection rounds
      // then scan shared memory instead of re-deriving 22 400 distances from global memory nine times
      __syncthreads();
      for (int p = threadIdx.x; p < P; p += ASSIGN_THREADS) {
        const int h = p / lv.W[l], w = p % lv.W[l];
        float d = __int_as_float(0x7f800000);
        if (h < vh && w < vw) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(bx + static_cast<long long>(lv.off[l] + p) * 4));
          const float cx = __fdiv_rn(__fadd_rn(a.x, a.z), 2.f), cy = __fdiv_rn(__fadd_rn(a.y, a.w), 2.f);
          const float dx = __fadd_rn(cx, -gcx), dy = __fadd_rn(cy, -gcy);
          d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        }
        dist_s[p] = d;
      }
      __syncthreads();
    }
    for (int r = 0; r < k_l; ++r) {
      MinPair best{3.0e38f, 0x7fffffff};
      if (staged) {
        for (int p = threadIdx.x; p < P; p += ASSIGN_THREADS) {
          const float d = dist_s[p];
          const int idx = lv.off[l] + p;
          if (!(d < 3.0e38f) || !lex_less(pd, pi, d, idx)) continue;
          if (lex_less(d, idx, best.d, best.i)) { best.d = d; best.i = idx; }
        }
      } else {
        for (int p = threadIdx.x; p < P; p += ASSIGN_THREADS) {
          const int h = p / lv.W[l], w = p % lv.W[l];
          if (h >= vh || w >= vw) continue;
          const float* a = bx + static_cast<long long>(lv.off[l] + p) * 4;
          const float cx = __fdiv_rn(__fadd_rn(a[0], a[2]), 2.f), cy = __fdiv_rn(__fadd_rn(a[1], a[3]), 2.f);
          const float dx = __fadd_rn(cx, -gcx), dy = __fadd_rn(cy, -gcy);
          const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
          const int idx = lv.off[l] + p;
          if (!lex_less(pd, pi, d, idx)) continue;   // already taken (or before the cursor)
          if (lex_less(d, idx, best.d, best.i)) { best.d = d; best.i = idx; }
        }
      }
      best = block_min(best, red);
      if (best.i == 0x7fffffff) break;
      pd = best.d; pi = best.i;
      if (threadIdx.x == 0) {
        cand[ncand] = best.i;
        cand_iou[ncand] = iou_fp32(bx + static_cast<long long>(best.i) * 4, gb);
        ++ncand;
      }
    }
  }
  __syncthreads();
  const int nc = ncand;
  if (threadIdx.x == 0) {
    // mean: fp32 cascade in chunks of 16 rows (ATen's outer-reduction sum order); std: unbiased, fp64 accumulation
    float lvl1 = 0.f, lvl0 = 0.f;
    for (int i = 0; i < nc; ++i) {
      lvl0 = __fadd_rn(lvl0, cand_iou[i]);
      if (((i + 1) & 15) == 0) { lvl1 = __fadd_rn(lvl1, lvl0); lvl0 = 0.f; }
    }
    const float mean = __fdiv_rn(__fadd_rn(lvl0, lvl1), static_cast<float>(nc));
    double m = 0.0, m2 = 0.0;
    for (int i = 0; i < nc; ++i) {
      const double x = static_cast<double>(cand_iou[i]);
      const double dlt = x - m;
      m += dlt / static_cast<double>(i + 1);
      m2 += dlt * (x - m);
    }
    const float sd = nc > 1 ? static_cast<float>(sqrt(m2 / static_cast<double>(nc - 1))) : nanf("");
    thr_s = __fadd_rn(mean, sd);
  }
  __syncthreads();
  const float thr = thr_s;
  for (int c = threadIdx.x; c < nc; c += ASSIGN_THREADS) {
    const int idx = cand[c];
    const float iou = cand_iou[c];
    if (!(iou >= thr)) continue;
    const float* a = bx + static_cast<long long>(idx) * 4;
    const float cx = __fdiv_rn(__fadd_rn(a[0], a[2]), 2.f), cy = __fdiv_rn(__fadd_rn(a[1], a[3]), 2.f);
    const float l_ = __fadd_rn(cx, -gb[0]), t_ = __fadd_rn(cy, -gb[1]);
    const float r_ = __fadd_rn(gb[2], -cx), b_ = __fadd_rn(gb[3], -cy);
    if (!(fminf(fminf(l_, t_), fminf(r_, b_)) > 0.01f)) continue;
    const unsigned long long key =
        (static_cast<unsigned long long>(__float_as_uint(iou)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - g);
    atomicMax(best_key + static_cast<long long>(b) * lv.total + idx, key);
  }
}

// per point: decode the ATSS winner.  grid over B*total.
__global__ void atss_resolve_kernel(const unsigned long long* __restrict__ best_key, long long n,
                                    int* __restrict__ assign, float* __restrict__ max_overlaps) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = best_key[i];
  if (k == 0ull) {
    assign[i] = -1;
    if (max_overlaps) max_overlaps[i] = -kInf;
  } else {
    assign[i] = static_cast<int>(0xFFFFFFFFu - static_cast<unsigned int>(k & 0xFFFFFFFFull));
    if (max_overlaps) max_overlaps[i] = __uint_as_float(static_cast<unsigned int>(k >> 32));
  }
}

// ------------------------------------------------------------------------------------------------ targets
// labels (background = num_classes on valid negatives AND on invalid points, lsnet_head.py:837,895),
// label_weights (1 on valid points, 0 on unmapped), per-image positive count.
__global__ void targets_kernel(const Levels lv, const int* __restrict__ valid_hw, const int* __restrict__ assign,
                               const int* __restrict__ gt_labels, int Gmax, int B, int num_classes,
                               int* __restrict__ labels, float* __restrict__ label_weights,
                               int* __restrict__ num_pos) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * lv.total) return;
  const int b = static_cast<int>(i / lv.total), p = static_cast<int>(i % lv.total);
  int l = 0;
  while (l + 1 < lv.n && p >= lv.off[l + 1]) ++l;
  const int q = p - lv.off[l];
  const bool valid = (q / lv.W[l] < valid_hw[(b * lv.n + l) * 2]) && (q % lv.W[l] < valid_hw[(b * lv.n + l) * 2 + 1]);
  const int a = valid ? assign[i] : -1;
  int lab = num_classes;
  if (a >= 0) {
    lab = gt_labels ? gt_labels[b * Gmax + a] : 1;
    atomicAdd(num_pos + b, 1);
  }
  if (labels) labels[i] = lab;
  if (label_weights) label_weights[i] = valid ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------------ pred boxes
// pred: NHWC [B, Hl, Wl, ldp] with D = 4*NP softplus'd slots [y-,y+,x-,x+] per landmark.  polygon == 0: box from the
// 4 extreme points; polygon == 1: min/max over the NP-1 contour points (centre excluded).
__global__ void pred_boxes_kernel(const float* __restrict__ pred, long long ldp, int NP, int polygon, int B, int Hl,
                                  int Wl, float stride, int level_off, int total, float* __restrict__ boxes) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int Pl = Hl * Wl;
  if (n >= B * Pl) return;
  const int b = n / Pl, pix = n % Pl;
  const float* p = pred + static_cast<long long>(n) * ldp;
  auto sgn = [&](int j, int xy) {   // xy: 0 = y pair, 1 = x pair
    const float pm = p[4 * j + 2 * xy], pp = p[4 * j + 2 * xy + 1];
    return pp > pm ? pp : -pm;
  };
  float x1, y1, x2, y2;
  if (!polygon) {
    x1 = sgn(1, 1); y1 = sgn(0, 0); x2 = sgn(3, 1); y2 = sgn(2, 0);
  } else {
    x1 = x2 = sgn(0, 1); y1 = y2 = sgn(0, 0);
    for (int j = 1; j < NP - 1; ++j) {
      const float x = sgn(j, 1), y = sgn(j, 0);
      x1 = fminf(x1, x); x2 = fmaxf(x2, x); y1 = fminf(y1, y); y2 = fmaxf(y2, y);
    }
  }
  const float ax = static_cast<float>(pix % Wl) * stride, ay = static_cast<float>(pix / Wl) * stride;
  float* o = boxes + (static_cast<long long>(b) * total + level_off + pix) * 4;
  o[0] = __fadd_rn(ax, __fmul_rn(x1, stride));
  o[1] = __fadd_rn(ay, __fmul_rn(y1, stride));
  o[2] = __fadd_rn(ax, __fmul_rn(x2, stride));
  o[3] = __fadd_rn(ay, __fmul_rn(y2, stride));
}

static int make_levels(Levels* lv, int n, const int* H, const int* W, const float* stride) {
  if (n < 1 || n > kMaxLevels) return set_error("assignment: %d pyramid levels unsupported (1..%d)", n, kMaxLevels);
  lv->n = n;
  int off = 0;
  for (int l = 0; l < n; ++l) {
    lv->H[l] = H[l]; lv->W[l] = W[l]; lv->stride[l] = stride[l]; lv->off[l] = off;
    off += H[l] * W[l];
  }
  lv->total = off;
  return 0;
}

}  // namespace lsn

using namespace lsn;

extern "C" int lsnet_centroid_assign(int num_levels, const int* level_h, const int* level_w, const float* level_stride,
                                     const int* valid_hw, const float* gt_bbox, const int* gt_count, int B, int Gmax,
                                     float scale, float* ws_best_d, int* ws_best_i, int* assign, void* stream) {
  Levels lv;
  if (int rc = make_levels(&lv, num_levels, level_h, level_w, level_stride)) return rc;
  if (B <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(assign, 0xFF, sizeof(int) * static_cast<size_t>(B) * lv.total, st);
  if (Gmax > 0) {
    centroid_nearest_kernel<<<dim3(Gmax, B), ASSIGN_THREADS, 0, st>>>(lv, valid_hw, gt_bbox, gt_count, Gmax, scale,
                                                                      ws_best_d, ws_best_i);
    if (int rc = check_launch("centroid_nearest")) return rc;
    centroid_resolve_kernel<<<B, 128, 0, st>>>(ws_best_d, ws_best_i, gt_count, Gmax, assign, lv.total);
    if (int rc = check_launch("centroid_resolve")) return rc;
  }
  return 0;
}

extern "C" int lsnet_atss_assign(int num_levels, const int* level_h, const int* level_w, const float* level_stride,
                                 const int* valid_hw, const float* boxes, const float* gt_bbox, const int* gt_count,
                                 int B, int Gmax, int topk, unsigned long long* ws_keys, int* assign,
                                 float* max_overlaps, void* stream) {
  Levels lv;
  if (int rc = make_levels(&lv, num_levels, level_h, level_w, level_stride)) return rc;
  if (B <= 0) return 0;
  if (topk > 16) return set_error("lsnet_atss_assign: topk=%d > 16 unsupported", topk);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = static_cast<long long>(B) * lv.total;
  cudaMemsetAsync(ws_keys, 0, sizeof(unsigned long long) * static_cast<size_t>(n), st);
  if (Gmax > 0) {
    int pmax = 0;
    for (int l = 0; l < lv.n; ++l) pmax = lv.H[l] * lv.W[l] > pmax ? lv.H[l] * lv.W[l] : pmax;
    static int max_smem = -1;
    if (max_smem < 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      cudaFuncSetAttribute(atss_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - 4096);
    }
    // boxes rows are 16-byte aligned (cudaMalloc'ed [B, total, 4] fp32): float4 loads in the staging pass
    const int cap = (static_cast<size_t>(pmax) * 4 <= static_cast<size_t>(max_smem - 4096) &&
                     reinterpret_cast<uintptr_t>(boxes) % 16 == 0) ? pmax : 0;
    atss_candidates_kernel<<<dim3(Gmax, B), ASSIGN_THREADS, static_cast<size_t>(cap) * 4, st>>>(
        lv, valid_hw, boxes, gt_bbox, gt_count, Gmax, topk, ws_keys, cap);
    if (int rc = check_launch("atss_candidates")) return rc;
  }
  atss_resolve_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(ws_keys, n, assign, max_overlaps);
  return check_launch("atss_resolve");
}

extern "C" int lsnet_assign_targets(int num_levels, const int* level_h, const int* level_w, const float* level_stride,
                                    const int* valid_hw, const int* assign, const int* gt_labels, int B, int Gmax,
                                    int num_classes, int* labels, float* label_weights, int* num_pos, void* stream) {
  Levels lv;
  if (int rc = make_levels(&lv, num_levels, level_h, level_w, level_stride)) return rc;
  if (B <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(num_pos, 0, sizeof(int) * B, st);
  const long long n = static_cast<long long>(B) * lv.total;
  targets_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(lv, valid_hw, assign, gt_labels, Gmax, B,
                                                                    num_classes, labels, label_weights, num_pos);
  return check_launch("targets");
}

extern "C" int lsnet_pred_boxes(const float* pred, long long ldp, int NP, int polygon, int B, int Hl, int Wl,
                                float stride, int level_off, int total_points, float* boxes, void* stream) {
  const int n = B * Hl * Wl;
  if (n <= 0) return 0;
  pred_boxes_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, ldp, NP, polygon, B, Hl, Wl,
                                                                                    stride, level_off, total_points,
                                                                                    boxes);
  return check_launch("pred_boxes");
}
