// Deformable-convolution sampling kernels (sm_100a): the bilinear-offset gather that builds the bf16 column
// matrix (forward / weight-gradient operand) and its adjoint (scatter of dCol to dX, reductions to dOffset/dMask).
//
// One code path serves DCNv1, DCNv2 (mask != null) and LSNet's pyramid DCN (scale_h/scale_w != 1, input extent
// (H,W) decoupled from the sampling grid (Ho,Wo)).  Arithmetic follows the reference kernels
//   im2col          mmdet/ops/dcn/src/cuda/deform_conv_cuda_kernel.cu:190-297, 847-910
//   col2im          :333-448, 912-970        col2im_coord  :486-615, 972-1044
// but the layout is B200-first: NHWC bf16 activations (one pixel = one contiguous channel vector, so every
// bilinear corner is a coalesced 16 B/lane load), pixel-major fp32 offsets/masks, and a [pixels, taps*C] bf16
// column matrix (tap-major, channel-minor) that the tcgen05 GEMM consumes K-major through TMA.
// HBM-bound: algorithmic bytes per output pixel = 2C (x) + 4*3*taps (offset,mask) + 2*taps*C (columns).
#include "common.cuh"
#include <stdlib.h>

#include "lsnet_internal.h"

namespace lsn {

struct DcnGeom {
  int B, H, W, C;          // input (sampled) feature map, NHWC
  int Ho, Wo;              // sampling / output grid
  int kh, kw, sh, sw, ph, pw, dh, dw;
  float scale_h, scale_w;  // pyramid: base grid is scaled, the learned offset is not (…kernel.cu:281-282)
  int dg;                  // deformable groups
  long long ldx, ldo, ldm, ldcol;
};

constexpr int PATCH_H = 4, PATCH_W = 8;   // 32 output pixels per CTA: keeps the sampled rows L1-resident
constexpr int GATHER_THREADS = 256;

struct Corner {
  float w[4];       // bilinear weights (0 when the corner is outside)
  long long o[4];   // pixel offsets (elements / ldx) of the 4 corners (clamped in range)
  bool inside;
  float lh, lw;
  bool v[4];
};

__device__ __forceinline__ Corner make_corner(const DcnGeom& g, int b, float h, float w) {
  Corner c;
  c.inside = (h > -1.f) && (w > -1.f) && (h < static_cast<float>(g.H)) && (w < static_cast<float>(g.W));
  const float hf = floorf(h), wf = floorf(w);
  const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
  const int h1 = h0 + 1, w1 = w0 + 1;
  const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
  c.lh = lh; c.lw = lw;
  c.v[0] = c.inside && h0 >= 0 && w0 >= 0;
  c.v[1] = c.inside && h0 >= 0 && w1 <= g.W - 1;
  c.v[2] = c.inside && h1 <= g.H - 1 && w0 >= 0;
  c.v[3] = c.inside && h1 <= g.H - 1 && w1 <= g.W - 1;
  c.w[0] = c.v[0] ? hh * hw : 0.f;
  c.w[1] = c.v[1] ? hh * lw : 0.f;
  c.w[2] = c.v[2] ? lh * hw : 0.f;
  c.w[3] = c.v[3] ? lh * lw : 0.f;
  const int ch0 = min(max(h0, 0), g.H - 1), ch1 = min(max(h1, 0), g.H - 1);
  const int cw0 = min(max(w0, 0), g.W - 1), cw1 = min(max(w1, 0), g.W - 1);
  const long long base = static_cast<long long>(b) * g.H;
  c.o[0] = ((base + ch0) * g.W + cw0) * g.ldx;
  c.o[1] = ((base + ch0) * g.W + cw1) * g.ldx;
  c.o[2] = ((base + ch1) * g.W + cw0) * g.ldx;
  c.o[3] = ((base + ch1) * g.W + cw1) * g.ldx;
  return c;
}

__device__ __forceinline__ void sample_pos(const DcnGeom& g, const float* __restrict__ off_px, int grp, int k,
                                           int ho, int wo, float* h, float* w) {
  const int i = k / g.kw, j = k % g.kw;
  const float oh = __ldg(off_px + grp * 2 * g.kh * g.kw + 2 * k);
  const float ow = __ldg(off_px + grp * 2 * g.kh * g.kw + 2 * k + 1);
  // mul then add, each rounded (matches the CPU oracle; the reference GPU build contracts this to one FMA)
  *h = __fadd_rn(__fmul_rn(static_cast<float>(ho * g.sh - g.ph + i * g.dh), g.scale_h), oh);
  *w = __fadd_rn(__fmul_rn(static_cast<float>(wo * g.sw - g.pw + j * g.dw), g.scale_w), ow);
}

// 16-byte streaming load that does not allocate in L1 (the column matrix is read exactly once; keep L1 for x)
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// grid: (patches_w, patches_h, B).  One warp per output pixel: lane k (< taps) derives the sampling position, the four
// (mask-folded) bilinear weights and the four corner pixel indices of tap k ONCE; the tap loop broadcasts them with
// shuffles, and every lane gathers/interpolates/stores the 8 channels it owns (16-byte accesses, coalesced per corner).
template <int U>   // taps in flight per lane: 4*U independent 16-byte gathers are issued before the first use
__global__ void __launch_bounds__(GATHER_THREADS)
dcn_im2col_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ offset,
                  const float* __restrict__ mask, __nv_bfloat16* __restrict__ col, const DcnGeom g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = GATHER_THREADS / 32;
  const int b = blockIdx.z;
  const int h_base = blockIdx.y * PATCH_H, w_base = blockIdx.x * PATCH_W;
  const int taps = g.kh * g.kw;
  const int cpg = g.C / g.dg;
  for (int pix = warp; pix < PATCH_H * PATCH_W; pix += nwarps) {
    const int ho = h_base + pix / PATCH_W, wo = w_base + pix % PATCH_W;
    if (ho >= g.Ho || wo >= g.Wo) continue;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    const float* off_px = offset + p * g.ldo;
    __nv_bfloat16* dst = col + p * g.ldcol;
    for (int grp = 0; grp < g.dg; ++grp) {
      for (int t0 = 0; t0 < taps; t0 += 32) {
        // ---- lane k: tap t0 + k ----
        float cw0 = 0.f, cw1 = 0.f, cw2 = 0.f, cw3 = 0.f;
        int co0 = 0, co1 = 0, co2 = 0, co3 = 0;
        if (t0 + lane < taps) {
          const int k = t0 + lane;
          float h, w;
          sample_pos(g, off_px, grp, k, ho, wo, &h, &w);
          const Corner cn = make_corner(g, b, h, w);
          float m = 1.f;
          if (mask) m = __ldg(mask + p * g.ldm + grp * taps + k);
          cw0 = cn.w[0] * m; cw1 = cn.w[1] * m; cw2 = cn.w[2] * m; cw3 = cn.w[3] * m;
          co0 = static_cast<int>(cn.o[0] / g.ldx); co1 = static_cast<int>(cn.o[1] / g.ldx);
          co2 = static_cast<int>(cn.o[2] / g.ldx); co3 = static_cast<int>(cn.o[3] / g.ldx);
        }
        const int nt = min(32, taps - t0);
        for (int kk = 0; kk < nt; kk += U) {
          float wq[U][4];
          long long oq[U][4];
          bool any[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int src = min(kk + u, nt - 1);
            wq[u][0] = __shfl_sync(0xffffffffu, cw0, src); wq[u][1] = __shfl_sync(0xffffffffu, cw1, src);
            wq[u][2] = __shfl_sync(0xffffffffu, cw2, src); wq[u][3] = __shfl_sync(0xffffffffu, cw3, src);
            oq[u][0] = static_cast<long long>(__shfl_sync(0xffffffffu, co0, src)) * g.ldx;
            oq[u][1] = static_cast<long long>(__shfl_sync(0xffffffffu, co1, src)) * g.ldx;
            oq[u][2] = static_cast<long long>(__shfl_sync(0xffffffffu, co2, src)) * g.ldx;
            oq[u][3] = static_cast<long long>(__shfl_sync(0xffffffffu, co3, src)) * g.ldx;
            any[u] = (kk + u < nt) &&
                     ((wq[u][0] != 0.f) || (wq[u][1] != 0.f) || (wq[u][2] != 0.f) || (wq[u][3] != 0.f));
          }
          for (int c0 = grp * cpg + lane * 8; c0 < (grp + 1) * cpg; c0 += 256) {
            uint4 ld[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (any[u]) {
#pragma unroll
                for (int q = 0; q < 4; ++q) ld[u][q] = __ldg(reinterpret_cast<const uint4*>(x + oq[u][q] + c0));
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (kk + u >= nt) continue;
              uint4 outv = make_uint4(0u, 0u, 0u, 0u);
              if (any[u]) {
                float f0[8], f1[8], f2[8], f3[8], acc[8];
                bf16x8_to_float(ld[u][0], f0); bf16x8_to_float(ld[u][1], f1);
                bf16x8_to_float(ld[u][2], f2); bf16x8_to_float(ld[u][3], f3);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  acc[e] = fmaf(wq[u][3], f3[e], fmaf(wq[u][2], f2[e], fmaf(wq[u][1], f1[e], wq[u][0] * f0[e])));
                outv = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                                  pack_bf16x2(acc[6], acc[7]));
              }
              st_stream(dst + static_cast<long long>(t0 + kk + u) * g.C + c0, outv);
            }
          }
        }
      }
    }
  }
}

// Adjoint.  gcol: [pixels, taps*C] bf16 (= dY . W).  One warp per output pixel; lane k derives tap k's sampling
// geometry once and the tap loop broadcasts it (as in the gather).  dOffset / dMask are channel reductions done with
// warp shuffles (no atomics).  dX is accumulated with 16-byte vector reds of packed bf16x2 (8 channels per lane per
// instruction): the kernel is bound by the SM-side RED issue rate (~1 cycle per lane), so halving the instruction count
// against fp32 v4 reds halves the time, and dX lands directly in the bf16 NHWC layout the upstream kernels consume.
// Corners whose bilinear weight is exactly zero (integer-aligned samples, e.g. zero-initialised conv_offset) are skipped.
// (Reference: atomics-only fp32 col2im, deform_conv_cuda_kernel.cu:346-388.)
__device__ __forceinline__ void red_bf16x8(__nv_bfloat16* dst, const float (&v)[8]) {
  const uint32_t a = pack_bf16x2(v[0], v[1]), b = pack_bf16x2(v[2], v[3]), c = pack_bf16x2(v[4], v[5]),
                 d = pack_bf16x2(v[6], v[7]);
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

__device__ __forceinline__ void red_f32x8(float* dst, const float (&v)[8]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
               : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

template <bool DX_FP32>
__global__ void __launch_bounds__(GATHER_THREADS, 3)
dcn_col2im_kernel(const __nv_bfloat16* __restrict__ gcol, const __nv_bfloat16* __restrict__ x,
                  const float* __restrict__ offset, const float* __restrict__ mask, void* __restrict__ dx,
                  float* __restrict__ doffset, float* __restrict__ dmask, const DcnGeom g, long long lddx,
                  long long lddo, long long lddm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = GATHER_THREADS / 32;
  const int b = blockIdx.z;
  const int h_base = blockIdx.y * PATCH_H, w_base = blockIdx.x * PATCH_W;
  const int taps = g.kh * g.kw;
  const int cpg = g.C / g.dg;
  for (int pix = warp; pix < PATCH_H * PATCH_W; pix += nwarps) {
    const int ho = h_base + pix / PATCH_W, wo = w_base + pix % PATCH_W;
    if (ho >= g.Ho || wo >= g.Wo) continue;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    const float* off_px = offset + p * g.ldo;
    for (int grp = 0; grp < g.dg; ++grp) {
      for (int t0 = 0; t0 < taps; t0 += 32) {
        // ---- lane k: geometry of tap t0 + k ----
        float lh_ = 0.f, lw_ = 0.f, m_ = 1.f;
        int vbits = 0, co0 = 0, co1 = 0, co2 = 0, co3 = 0;
        if (t0 + lane < taps) {
          const int k = t0 + lane;
          float h, w;
          sample_pos(g, off_px, grp, k, ho, wo, &h, &w);
          const Corner cn = make_corner(g, b, h, w);
          if (mask) m_ = __ldg(mask + p * g.ldm + grp * taps + k);
          lh_ = cn.lh; lw_ = cn.lw;
          vbits = (cn.v[0] ? 1 : 0) | (cn.v[1] ? 2 : 0) | (cn.v[2] ? 4 : 0) | (cn.v[3] ? 8 : 0) | (cn.inside ? 16 : 0);
          co0 = static_cast<int>(cn.o[0] / g.ldx); co1 = static_cast<int>(cn.o[1] / g.ldx);
          co2 = static_cast<int>(cn.o[2] / g.ldx); co3 = static_cast<int>(cn.o[3] / g.ldx);
        }
        const int nt = min(32, taps - t0);
        float my_gh = 0.f, my_gw = 0.f, my_gm = 0.f;   // results of tap (t0 + lane), filled by the reductions below
        for (int kk = 0; kk < nt; ++kk) {
          const float lh = __shfl_sync(0xffffffffu, lh_, kk), lw = __shfl_sync(0xffffffffu, lw_, kk);
          const float m = __shfl_sync(0xffffffffu, m_, kk);
          const int vb = __shfl_sync(0xffffffffu, vbits, kk);
          const int oi[4] = {__shfl_sync(0xffffffffu, co0, kk), __shfl_sync(0xffffffffu, co1, kk),
                             __shfl_sync(0xffffffffu, co2, kk), __shfl_sync(0xffffffffu, co3, kk)};
          float gh = 0.f, gw = 0.f, gm = 0.f;
          if (vb & 16) {
            const float hh = 1.f - lh, hw = 1.f - lw;
            const float wq[4] = {(vb & 1) ? hh * hw : 0.f, (vb & 2) ? hh * lw : 0.f, (vb & 4) ? lh * hw : 0.f,
                                 (vb & 8) ? lh * lw : 0.f};
            const __nv_bfloat16* src = gcol + p * g.ldcol + static_cast<long long>(t0 + kk) * g.C;
            for (int c0 = grp * cpg + lane * 8; c0 < (grp + 1) * cpg; c0 += 256) {
              float gc[8];
              bf16x8_to_float(ld_stream(src + c0), gc);
              float xv[4][8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (vb & (1 << q)) {
                  bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(oi[q]) * g.ldx + c0)), xv[q]);
                } else {
#pragma unroll
                  for (int e = 0; e < 8; ++e) xv[q][e] = 0.f;
                }
              }
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float val = wq[0] * xv[0][e] + wq[1] * xv[1][e] + wq[2] * xv[2][e] + wq[3] * xv[3][e];
                // d(val)/dh, d(val)/dw: get_coordinate_weight (…kernel.cu:145-188), out-of-range corners dropped
                const float dvh = -hw * xv[0][e] - lw * xv[1][e] + hw * xv[2][e] + lw * xv[3][e];
                const float dvw = -hh * xv[0][e] + hh * xv[1][e] - lh * xv[2][e] + lh * xv[3][e];
                gm = fmaf(gc[e], val, gm);
                gh = fmaf(gc[e] * m, dvh, gh);
                gw = fmaf(gc[e] * m, dvw, gw);
              }
              if (dx) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float sc = wq[q] * m;
                  if (sc == 0.f) continue;
                  float v[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = sc * gc[e];
                  if (DX_FP32) red_f32x8(static_cast<float*>(dx) + static_cast<long long>(oi[q]) * lddx + c0, v);
                  else red_bf16x8(static_cast<__nv_bfloat16*>(dx) + static_cast<long long>(oi[q]) * lddx + c0, v);
                }
              }
            }
          }
          gh = warp_sum(gh); gw = warp_sum(gw); gm = warp_sum(gm);
          if (lane == kk) { my_gh = gh; my_gw = gw; my_gm = gm; }
        }
        if (t0 + lane < taps) {
          const int k = t0 + lane;
          doffset[p * lddo + grp * 2 * taps + 2 * k] = my_gh;
          doffset[p * lddo + grp * 2 * taps + 2 * k + 1] = my_gw;
          if (dmask) dmask[p * lddm + grp * taps + k] = my_gm;
        }
      }
    }
  }
}

static int check_geom(const char* who, int C, int dg, long long ldx, long long ldcol) {
  if (dg < 1 || C % dg || (C / dg) % 8 || (ldx % 8) || (ldcol % 8))
    return set_error("%s: need C/deformable_groups %% 8 == 0 and 16-byte aligned pitches (C=%d dg=%d)", who, C, dg);
  return 0;
}

}  // namespace lsn

using namespace lsn;

extern "C" int lsnet_dcn_im2col_bf16(const void* x, int B, int H, int W, int C, long long ldx, const float* offset,
                                     long long ldo, const float* mask, long long ldm, int Ho, int Wo, int kh, int kw,
                                     int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                     float scale_h, float scale_w, int deformable_groups, void* col, long long ldcol,
                                     void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if (int rc = check_geom("lsnet_dcn_im2col_bf16", C, deformable_groups, ldx, ldcol)) return rc;
  DcnGeom g{B, H, W, C, Ho, Wo, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, scale_h, scale_w,
            deformable_groups, ldx, ldo, ldm, ldcol};
  dim3 grid((Wo + PATCH_W - 1) / PATCH_W, (Ho + PATCH_H - 1) / PATCH_H, B);
  const double taps = kh * kw, px = static_cast<double>(B) * Ho * Wo;
  // algorithmic bytes: x read once (B*H*W*2C) + offsets/mask (4*(2+[mask])*taps per px) + bf16 columns written
  const double bytes = static_cast<double>(B) * H * W * 2.0 * C + px * 4.0 * taps * (mask ? 3 : 2) + px * 2.0 * taps * C;
  const int th = timing_begin(TC_IM2COL, bytes, static_cast<cudaStream_t>(stream));
  static int unroll = 0;
  if (!unroll) {
    const char* e = getenv("LSNET_IM2COL_U");
    unroll = e ? atoi(e) : 1;
    if (unroll < 1 || unroll > 3) unroll = 1;
  }
  cudaStream_t st_ = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* cp = static_cast<__nv_bfloat16*>(col);
  if (unroll == 1) dcn_im2col_kernel<1><<<grid, GATHER_THREADS, 0, st_>>>(xp, offset, mask, cp, g);
  else if (unroll == 2) dcn_im2col_kernel<2><<<grid, GATHER_THREADS, 0, st_>>>(xp, offset, mask, cp, g);
  else dcn_im2col_kernel<3><<<grid, GATHER_THREADS, 0, st_>>>(xp, offset, mask, cp, g);
  timing_end(th, static_cast<cudaStream_t>(stream));
  return check_launch("dcn_im2col");
}

extern "C" int lsnet_dcn_col2im_bf16(const void* gcol, long long ldcol, const void* x, int B, int H, int W, int C,
                                     long long ldx, const float* offset, long long ldo, const float* mask,
                                     long long ldm, int Ho, int Wo, int kh, int kw, int stride_h, int stride_w,
                                     int pad_h, int pad_w, int dil_h, int dil_w, float scale_h, float scale_w,
                                     int deformable_groups, void* dx, long long lddx, int dx_fp32, float* doffset,
                                     long long lddo, float* dmask, long long lddm, void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if (int rc = check_geom("lsnet_dcn_col2im_bf16", C, deformable_groups, ldx, ldcol)) return rc;
  if (dx && (lddx % 8)) return set_error("lsnet_dcn_col2im_bf16: dx pitch must be a multiple of 8 bf16");
  DcnGeom g{B, H, W, C, Ho, Wo, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, scale_h, scale_w,
            deformable_groups, ldx, ldo, ldm, ldcol};
  dim3 grid((Wo + PATCH_W - 1) / PATCH_W, (Ho + PATCH_H - 1) / PATCH_H, B);
  const double taps = kh * kw, px = static_cast<double>(B) * Ho * Wo;
  // algorithmic bytes: dCol read (2*taps*C per px) + x read + offsets/mask read + dX written (bf16) + dOffset/dMask
  const double bytes = px * 2.0 * taps * C + static_cast<double>(B) * H * W * (2.0 * C + (dx ? (dx_fp32 ? 4.0 : 2.0) * C : 0.0)) +
                       px * 4.0 * taps * (mask ? 3 : 2) * 2.0;
  const int th = timing_begin(TC_COL2IM, bytes, static_cast<cudaStream_t>(stream));
  if (dx_fp32)
    dcn_col2im_kernel<true><<<grid, GATHER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(gcol), static_cast<const __nv_bfloat16*>(x), offset, mask, dx, doffset, dmask,
        g, lddx, lddo, lddm);
  else
    dcn_col2im_kernel<false><<<grid, GATHER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(gcol), static_cast<const __nv_bfloat16*>(x), offset, mask, dx, doffset, dmask,
        g, lddx, lddo, lddm);
  timing_end(th, static_cast<cudaStream_t>(stream));
  return check_launch("dcn_col2im");
}
