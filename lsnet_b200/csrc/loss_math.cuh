// Row-level arithmetic of the LSNet losses, written __host__ __device__ so the exact code the kernels run can
// also be exercised on the CPU-only build box (tests/test_loss_math_host.py via lsnet_host_* entry points).
//
// cross_iou_row follows mmdet/models/losses/cross_iou_loss.py:61-132 (bbox / polygon / keypoint), including the
// un-detached v^2/(1-IoU+v) term; gradients follow torch's autograd conventions for the ops the reference uses
// (max/min over a stacked pair return the first index on ties; clamp passes gradient on the closed range).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LSN_HD __host__ __device__ __forceinline__
#else
#define LSN_HD inline
#endif

namespace lsn {

enum { LOSS_BBOX = 0, LOSS_POLYGON = 1, LOSS_KEYPOINT = 2 };
constexpr int kMaxD = 160;   // 4*(36+1) = 148 for segm

struct BoxPenalty {
  float value;          // rho2/c2 + v^2/(1-IoU+v)
  float d_iou;          // d(value)/d(IoU)
  float d_box[4];       // d(value)/d(x1,y1,x2,y2) of the predicted box
};

// CIoU-style penalty of cross_iou_loss.py:104-128.  iou enters only through v^2/(1-iou+v).
LSN_HD BoxPenalty box_penalty(const float b[4], const float g[4], float iou, float eps) {
  BoxPenalty r;
  const float ex1 = fminf(b[0], g[0]), ey1 = fminf(b[1], g[1]);
  const float ex2 = fmaxf(b[2], g[2]), ey2 = fmaxf(b[3], g[3]);
  const float cw = fmaxf(ex2 - ex1, 0.f), ch = fmaxf(ey2 - ey1, 0.f);
  const float c2 = cw * cw + ch * ch + eps;
  const float w1 = b[2] - b[0], h1 = b[3] - b[1] + eps;
  const float w2 = g[2] - g[0], h2 = g[3] - g[1] + eps;
  const float sx = (g[0] + g[2]) - (b[0] + b[2]);
  const float sy = (g[1] + g[3]) - (b[1] + b[3]);
  const float rho2 = sx * sx / 4.f + sy * sy / 4.f;
  const float factor = 4.f / (3.14159265358979323846f * 3.14159265358979323846f);
  const float da = atanf(w2 / h2) - atanf(w1 / h1);
  const float v = factor * da * da;
  const float A = 1.f - iou + v;
  r.value = rho2 / c2 + v * v / A;
  // d(v^2/A): A = 1 - iou + v
  const float dterm_dv = (2.f * v * A - v * v) / (A * A);
  r.d_iou = v * v / (A * A);
  // dv/d(box): v = f*da^2, da = atan(w2/h2) - atan(w1/h1)
  const float den = h1 * h1 + w1 * w1;
  const float datan_dw1 = h1 / den, datan_dh1 = -w1 / den;
  const float dv_dw1 = -2.f * factor * da * datan_dw1;
  const float dv_dh1 = -2.f * factor * da * datan_dh1;
  // enclosing box: torch.min/max (binary) split the gradient evenly on exact ties
  const float m_x1 = b[0] < g[0] ? 1.f : (b[0] == g[0] ? 0.5f : 0.f);
  const float m_y1 = b[1] < g[1] ? 1.f : (b[1] == g[1] ? 0.5f : 0.f);
  const float m_x2 = b[2] > g[2] ? 1.f : (b[2] == g[2] ? 0.5f : 0.f);
  const float m_y2 = b[3] > g[3] ? 1.f : (b[3] == g[3] ? 0.5f : 0.f);
  const float cw_on = (ex2 - ex1) >= 0.f ? 1.f : 0.f, ch_on = (ey2 - ey1) >= 0.f ? 1.f : 0.f;
  const float dc2_dx1 = 2.f * cw * (-m_x1) * cw_on, dc2_dx2 = 2.f * cw * m_x2 * cw_on;
  const float dc2_dy1 = 2.f * ch * (-m_y1) * ch_on, dc2_dy2 = 2.f * ch * m_y2 * ch_on;
  const float k = rho2 / (c2 * c2);
  r.d_box[0] = (-sx / 2.f) / c2 - k * dc2_dx1 + dterm_dv * (-dv_dw1);
  r.d_box[2] = (-sx / 2.f) / c2 - k * dc2_dx2 + dterm_dv * (dv_dw1);
  r.d_box[1] = (-sy / 2.f) / c2 - k * dc2_dy1 + dterm_dv * (-dv_dh1);
  r.d_box[3] = (-sy / 2.f) / c2 - k * dc2_dy2 + dterm_dv * (dv_dh1);
  return r;
}

// signed coordinate of a (-,+) slot pair: torch.max over the pair, index 0 (the '-' slot) wins ties and is negated
// (cross_iou_loss.py:11-14).  *which = slot that carries the gradient, *sign = its derivative.
LSN_HD float signed_coord(float pm, float pp, int* which, float* sign) {
  if (pp > pm) { *which = 1; *sign = 1.f; return pp; }
  *which = 0; *sign = -1.f; return -pm;
}

// One row.  pred/target: D = 4*(L+1) slots laid out per landmark as [y-, y+, x-, x+]; sel marks the "true" slot of
// each pair; anchor = (x, y); vs: L visibility flags (keypoint only).  Returns the un-weighted row loss; if grad
// != nullptr writes d(loss)/d(pred[d]).
LSN_HD float cross_iou_row(int type, const float* pred, const float* target, const uint8_t* sel, int D,
                           const float* anchor, const float* bbox_gt, const float* vs, float eps, float alpha,
                           int pstride, float* grad) {
  float t[kMaxD];
  for (int d = 0; d < D; ++d) t[d] = sel[d] ? target[d] : alpha * target[d ^ 1];   // :65-66
  if (grad) for (int d = 0; d < D; ++d) grad[d] = 0.f;

  if (type == LOSS_KEYPOINT) {
    const int pairs = D / 2, L = D / 4 - 1;
    float acc = 0.f;
    for (int q = 0; q < pairs; ++q) {
      const float p0 = pred[2 * q], p1 = pred[2 * q + 1], t0 = t[2 * q], t1 = t[2 * q + 1];
      const float mx0 = fmaxf(fmaxf(p0, t0), eps), mx1 = fmaxf(fmaxf(p1, t1), eps);
      const float mn0 = fminf(p0, t0), mn1 = fminf(p1, t1);
      const float smin = mn0 + mn1, smax = mx0 + mx1;
      float vis = 1.f;
      if (q < 2 * L) vis = vs[q / 2] > 0.f ? 1.f : 0.f;   // :92-94 (centre always counted)
      acc += vis * smin / smax;
      if (grad) {
        // loss = 1 - sum_q vis*IoU_q / pairs
        const float s = -vis / static_cast<float>(pairs);
        const float dmin0 = p0 <= t0 ? 1.f : 0.f, dmin1 = p1 <= t1 ? 1.f : 0.f;
        const float dmax0 = (p0 >= t0 && p0 >= eps) ? 1.f : 0.f, dmax1 = (p1 >= t1 && p1 >= eps) ? 1.f : 0.f;
        grad[2 * q] = s * (dmin0 * smax - smin * dmax0) / (smax * smax);
        grad[2 * q + 1] = s * (dmin1 * smax - smin * dmax1) / (smax * smax);
      }
    }
    return 1.f - acc / static_cast<float>(pairs);
  }

  float iou;
  float b[4];
  int bslot[4];
  float bsign[4];
  if (type == LOSS_BBOX) {
    float smin = 0.f, smax = 0.f;
    for (int d = 0; d < D; ++d) { smin += fminf(pred[d], t[d]); smax += fmaxf(pred[d], t[d]); }
    iou = smin / smax;
    if (grad) {
      for (int d = 0; d < D; ++d) {
        const float dmin = pred[d] <= t[d] ? 1.f : 0.f, dmax = pred[d] >= t[d] ? 1.f : 0.f;
        grad[d] = (dmin * smax - smin * dmax) / (smax * smax);   // d(iou)/d(pred), sign applied below
      }
    }
    // box from the 4 extreme points (:10-33): left = x of point 1, top = y of point 0, right = x of 3, bottom = y of 2
    int wh; float sg;
    b[0] = signed_coord(pred[4 * 1 + 2], pred[4 * 1 + 3], &wh, &sg) + anchor[0]; bslot[0] = 4 * 1 + 2 + wh; bsign[0] = sg;
    b[1] = signed_coord(pred[4 * 0 + 0], pred[4 * 0 + 1], &wh, &sg) + anchor[1]; bslot[1] = 4 * 0 + 0 + wh; bsign[1] = sg;
    b[2] = signed_coord(pred[4 * 3 + 2], pred[4 * 3 + 3], &wh, &sg) + anchor[0]; bslot[2] = 4 * 3 + 2 + wh; bsign[2] = sg;
    b[3] = signed_coord(pred[4 * 2 + 0], pred[4 * 2 + 1], &wh, &sg) + anchor[1]; bslot[3] = 4 * 2 + 0 + wh; bsign[3] = sg;
  } else {   // polygon (:68-77, :35-59)
    const int npts = D / 4;            // landmarks + centre
    float acc = 0.f;
    float gmin[16], gmax[16];
    for (int i = 0; i < pstride; ++i) {
      float smin = 0.f, smax = 0.f;
      for (int j = i; j < npts; j += pstride)
        for (int s = 0; s < 4; ++s) {
          smin += fminf(pred[4 * j + s], t[4 * j + s]);
          smax += fmaxf(pred[4 * j + s], t[4 * j + s]);
        }
      gmin[i] = smin; gmax[i] = smax;
      acc += smin / smax;
    }
    iou = acc / static_cast<float>(pstride);
    if (grad) {
      for (int j = 0; j < npts; ++j) {
        const int i = j % pstride;
        for (int s = 0; s < 4; ++s) {
          const int d = 4 * j + s;
          const float dmin = pred[d] <= t[d] ? 1.f : 0.f, dmax = pred[d] >= t[d] ? 1.f : 0.f;
          grad[d] = (dmin * gmax[i] - gmin[i] * dmax) / (gmax[i] * gmax[i]) / static_cast<float>(pstride);
        }
      }
    }
    // box = min/max over the contour points (centre excluded), first index wins ties
    const int np = npts - 1;
    float xmin = 0.f, xmax = 0.f, ymin = 0.f, ymax = 0.f;
    for (int j = 0; j < np; ++j) {
      int wy, wx; float sy_, sx_;
      const float y = signed_coord(pred[4 * j], pred[4 * j + 1], &wy, &sy_) + anchor[1];
      const float x = signed_coord(pred[4 * j + 2], pred[4 * j + 3], &wx, &sx_) + anchor[0];
      if (j == 0 || x < xmin) { xmin = x; bslot[0] = 4 * j + 2 + wx; bsign[0] = sx_; }
      if (j == 0 || y < ymin) { ymin = y; bslot[1] = 4 * j + wy; bsign[1] = sy_; }
      if (j == 0 || x > xmax) { xmax = x; bslot[2] = 4 * j + 2 + wx; bsign[2] = sx_; }
      if (j == 0 || y > ymax) { ymax = y; bslot[3] = 4 * j + wy; bsign[3] = sy_; }
    }
    b[0] = xmin; b[1] = ymin; b[2] = xmax; b[3] = ymax;
  }
  const BoxPenalty pen = box_penalty(b, bbox_gt, iou, eps);
  if (grad) {
    // loss = 1 - iou + pen(box, iou)
    const float diou = -1.f + pen.d_iou;
    for (int d = 0; d < D; ++d) grad[d] *= diou;
    for (int e = 0; e < 4; ++e) grad[bslot[e]] += pen.d_box[e] * bsign[e];
  }
  return 1.f - (iou - pen.value);
}

// Directional regression targets of LSHead.get_bbox_gt_reg / get_poly_gt_reg (lsnet_head.py:402-454) for one row.
// gt: (x,y) pairs of the NP landmarks (incl. centre); anchor (x,y); positive = row weight > 0.
// Writes target[4*j + {0,1,2,3}] = [y-, y+, x-, x+] and the boolean slot mask sel (computed from the offsets even
// on negative rows, where gt is all-zero and the targets are zeroed).
LSN_HD void directional_target_row(const float* gt, int NP, const float* anchor, bool positive, float* target,
                                   uint8_t* sel) {
  for (int j = 0; j < NP; ++j) {
    const float ox = gt[2 * j] - anchor[0], oy = gt[2 * j + 1] - anchor[1];
    const bool xpos = ox >= 0.f, ypos = oy >= 0.f;
    const float ax = positive ? fabsf(ox) : 0.f, ay = positive ? fabsf(oy) : 0.f;
    target[4 * j + 0] = ypos ? 0.f : ay;  sel[4 * j + 0] = !ypos;
    target[4 * j + 1] = ypos ? ay : 0.f;  sel[4 * j + 1] = ypos;
    target[4 * j + 2] = xpos ? 0.f : ax;  sel[4 * j + 2] = !xpos;
    target[4 * j + 3] = xpos ? ax : 0.f;  sel[4 * j + 3] = xpos;
  }
}

// Sigmoid focal loss element (sigmoid_focal_loss_cuda.cu:23-97): label t, class d.
LSN_HD float focal_elem(float x, int t, int d, float gamma, float alpha, float* grad) {
  const float c1 = (t == d) ? 1.f : 0.f;
  const float c2 = (t >= 0 && t != d) ? 1.f : 0.f;
  const float zn = 1.f - alpha, zp = alpha;
  const float p = 1.f / (1.f + expf(-x));
  const float FLT_MIN_ = 1.17549435e-38f;
  const float logp = logf(fmaxf(p, FLT_MIN_));
  const float ge = x >= 0.f ? 1.f : 0.f;
  const float log1mp = -1.f * x * ge - logf(1.f + expf(x - 2.f * x * ge));
  const float term1 = powf(1.f - p, gamma) * logp;
  const float term2 = powf(p, gamma) * log1mp;
  if (grad) {
    const float g1 = powf(1.f - p, gamma) * (1.f - p - (p * gamma * logp));
    const float g2 = powf(p, gamma) * (log1mp * (1.f - p) * gamma - p);
    *grad = -c1 * g1 * zp - c2 * g2 * zn;
  }
  return -c1 * term1 * zp - c2 * term2 * zn;
}

}  // namespace lsn
