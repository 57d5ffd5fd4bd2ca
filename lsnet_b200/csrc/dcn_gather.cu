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

__device__ __forceinline__ void bf16x8_to_float(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

// grid: (patches_w, patches_h, B).  Each warp walks (pixel, tap) items of the patch; lanes own 8 channels each.
__global__ void __launch_bounds__(GATHER_THREADS)
dcn_im2col_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ offset,
                  const float* __restrict__ mask, __nv_bfloat16* __restrict__ col, const DcnGeom g) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = GATHER_THREADS / 32;
  const int b = blockIdx.z;
  const int h_base = blockIdx.y * PATCH_H, w_base = blockIdx.x * PATCH_W;
  const int taps = g.kh * g.kw;
  const int cpg = g.C / g.dg;
  const int items = PATCH_H * PATCH_W * taps;
  for (int item = warp; item < items; item += nwarps) {
    const int pix = item / taps, k = item % taps;
    const int ho = h_base + pix / PATCH_W, wo = w_base + pix % PATCH_W;
    if (ho >= g.Ho || wo >= g.Wo) continue;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    const float* off_px = offset + p * g.ldo;
    __nv_bfloat16* dst = col + p * g.ldcol + static_cast<long long>(k) * g.C;
    for (int c0 = lane * 8; c0 < g.C; c0 += 256) {
      const int grp = c0 / cpg;
      float h, w;
      sample_pos(g, off_px, grp, k, ho, wo, &h, &w);
      const Corner cn = make_corner(g, b, h, w);
      float m = 1.f;
      if (mask) m = __ldg(mask + p * g.ldm + grp * taps + k);
      uint4 outv = make_uint4(0u, 0u, 0u, 0u);
      if (cn.inside) {
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = __ldg(reinterpret_cast<const uint4*>(x + cn.o[q] + c0));
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
          bf16x8_to_float(u[q], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(cn.w[q], f[e], acc[e]);
        }
        outv = make_uint4(pack_bf16x2(acc[0] * m, acc[1] * m), pack_bf16x2(acc[2] * m, acc[3] * m),
                          pack_bf16x2(acc[4] * m, acc[5] * m), pack_bf16x2(acc[6] * m, acc[7] * m));
      }
      *reinterpret_cast<uint4*>(dst + c0) = outv;
    }
  }
}

// Adjoint.  gcol: [pixels, taps*C] bf16 (= dY . W).  dOffset / dMask are channel reductions done with warp shuffles
// (one warp owns a whole (pixel, tap) item, so no atomics there).  dX (fp32 NHWC) is accumulated in two tiers: a
// shared-memory window of WIN_H x WIN_W input pixels around the CTA's output patch absorbs every contribution that
// lands near the patch (all of them while |offset| < 1 at scale 1: ~36 contributions per input pixel collapse into
// one), and is flushed once with vector reds; contributions outside the window go straight to global reds.  This cuts
// L2 atomic traffic (the reference's col2im is atomics-only: deform_conv_cuda_kernel.cu:346-388) by up to ~12x.
constexpr int WIN_H = 8, WIN_W = 12;

template <bool USE_WIN>
__global__ void __launch_bounds__(GATHER_THREADS)
dcn_col2im_kernel(const __nv_bfloat16* __restrict__ gcol, const __nv_bfloat16* __restrict__ x,
                  const float* __restrict__ offset, const float* __restrict__ mask, float* __restrict__ dx,
                  float* __restrict__ doffset, float* __restrict__ dmask, const DcnGeom g, long long lddx,
                  long long lddo, long long lddm) {
  extern __shared__ float win[];   // [WIN_H*WIN_W][C], channel c stored at (c%8)*(C/8) + c/8 (bank-conflict free)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = GATHER_THREADS / 32;
  const int b = blockIdx.z;
  const int h_base = blockIdx.y * PATCH_H, w_base = blockIdx.x * PATCH_W;
  const int taps = g.kh * g.kw;
  const int cpg = g.C / g.dg;
  const int c8 = g.C / 8;
  // window origin: one pixel before the smallest un-offset sampling position of the patch
  const int wy0 = static_cast<int>(floorf(static_cast<float>(h_base * g.sh - g.ph) * g.scale_h)) - 1;
  const int wx0 = static_cast<int>(floorf(static_cast<float>(w_base * g.sw - g.pw) * g.scale_w)) - 1;
  if (USE_WIN && dx) {
    for (int i = threadIdx.x; i < WIN_H * WIN_W * g.C; i += GATHER_THREADS) win[i] = 0.f;
    __syncthreads();
  }
  const int items = PATCH_H * PATCH_W * taps * g.dg;
  for (int item = warp; item < items; item += nwarps) {
    const int grp = item % g.dg;
    const int k = (item / g.dg) % taps;
    const int pix = item / (g.dg * taps);
    const int ho = h_base + pix / PATCH_W, wo = w_base + pix % PATCH_W;
    if (ho >= g.Ho || wo >= g.Wo) continue;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    const float* off_px = offset + p * g.ldo;
    float h, w;
    sample_pos(g, off_px, grp, k, ho, wo, &h, &w);
    const Corner cn = make_corner(g, b, h, w);
    float m = 1.f;
    if (mask) m = __ldg(mask + p * g.ldm + grp * taps + k);
    float gh = 0.f, gw = 0.f, gm = 0.f;
    if (cn.inside) {
      const __nv_bfloat16* src = gcol + p * g.ldcol + static_cast<long long>(k) * g.C;
      const float hh = 1.f - cn.lh, hw = 1.f - cn.lw;
      // window slot of each corner (-1: outside the window -> global red)
      int wslot[4];
      {
        const int h0 = static_cast<int>(floorf(h)), w0 = static_cast<int>(floorf(w));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int yy = h0 + (q >> 1) - wy0, xx = w0 + (q & 1) - wx0;
          wslot[q] = (USE_WIN && yy >= 0 && yy < WIN_H && xx >= 0 && xx < WIN_W) ? yy * WIN_W + xx : -1;
        }
      }
      for (int c0 = grp * cpg + lane * 8; c0 < (grp + 1) * cpg; c0 += 256) {
        float gc[8];
        bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(src + c0)), gc);
        float xv[4][8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (cn.v[q]) {
            bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(x + cn.o[q] + c0)), xv[q]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) xv[q][e] = 0.f;
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float val = cn.w[0] * xv[0][e] + cn.w[1] * xv[1][e] + cn.w[2] * xv[2][e] + cn.w[3] * xv[3][e];
          // d(val)/dh and d(val)/dw: get_coordinate_weight (…kernel.cu:145-188), out-of-range corners dropped
          const float dvh = -hw * xv[0][e] - cn.lw * xv[1][e] + hw * xv[2][e] + cn.lw * xv[3][e];
          const float dvw = -hh * xv[0][e] + hh * xv[1][e] - cn.lh * xv[2][e] + cn.lh * xv[3][e];
          gm = fmaf(gc[e], val, gm);
          gh = fmaf(gc[e] * m, dvh, gh);
          gw = fmaf(gc[e] * m, dvw, gw);
        }
        if (dx) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (!cn.v[q]) continue;
            const float s = cn.w[q] * m;
            if (wslot[q] >= 0) {
              float* d = win + static_cast<long long>(wslot[q]) * g.C + (c0 >> 3);
#pragma unroll
              for (int e = 0; e < 8; ++e) atomicAdd(d + e * c8, s * gc[e]);
            } else {
              float* d = dx + (cn.o[q] / g.ldx) * lddx + c0;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(s * gc[0]), "f"(s * gc[1]),
                           "f"(s * gc[2]), "f"(s * gc[3])
                           : "memory");
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(s * gc[4]),
                           "f"(s * gc[5]), "f"(s * gc[6]), "f"(s * gc[7])
                           : "memory");
            }
          }
        }
      }
    }
    gh = warp_sum(gh);
    gw = warp_sum(gw);
    gm = warp_sum(gm);
    if (lane == 0) {
      doffset[p * lddo + grp * 2 * taps + 2 * k] = gh;
      doffset[p * lddo + grp * 2 * taps + 2 * k + 1] = gw;
      if (dmask) dmask[p * lddm + grp * taps + k] = gm;
    }
  }
  if (USE_WIN && dx) {
    __syncthreads();
    // flush: one thread per (window pixel, 8-channel vector); untouched vectors are skipped
    for (int i = threadIdx.x; i < WIN_H * WIN_W * c8; i += GATHER_THREADS) {
      const int slot = i / c8, v = i % c8;
      const int yy = wy0 + slot / WIN_W, xx = wx0 + slot % WIN_W;
      if (yy < 0 || yy >= g.H || xx < 0 || xx >= g.W) continue;
      const float* s = win + static_cast<long long>(slot) * g.C + v;
      float f[8];
      bool any = false;
#pragma unroll
      for (int e = 0; e < 8; ++e) { f[e] = s[e * c8]; any |= (f[e] != 0.f); }
      if (!any) continue;
      float* d = dx + ((static_cast<long long>(b) * g.H + yy) * g.W + xx) * lddx + v * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3])
                   : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(f[4]), "f"(f[5]), "f"(f[6]),
                   "f"(f[7])
                   : "memory");
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
  dcn_im2col_kernel<<<grid, GATHER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), offset, mask, static_cast<__nv_bfloat16*>(col), g);
  timing_end(th, static_cast<cudaStream_t>(stream));
  return check_launch("dcn_im2col");
}

extern "C" int lsnet_dcn_col2im_bf16(const void* gcol, long long ldcol, const void* x, int B, int H, int W, int C,
                                     long long ldx, const float* offset, long long ldo, const float* mask,
                                     long long ldm, int Ho, int Wo, int kh, int kw, int stride_h, int stride_w,
                                     int pad_h, int pad_w, int dil_h, int dil_w, float scale_h, float scale_w,
                                     int deformable_groups, float* dx, long long lddx, float* doffset, long long lddo,
                                     float* dmask, long long lddm, void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if (int rc = check_geom("lsnet_dcn_col2im_bf16", C, deformable_groups, ldx, ldcol)) return rc;
  if (dx && (lddx % 4)) return set_error("lsnet_dcn_col2im_bf16: dx pitch must be a multiple of 4 floats");
  DcnGeom g{B, H, W, C, Ho, Wo, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, scale_h, scale_w,
            deformable_groups, ldx, ldo, ldm, ldcol};
  dim3 grid((Wo + PATCH_W - 1) / PATCH_W, (Ho + PATCH_H - 1) / PATCH_H, B);
  const double taps = kh * kw, px = static_cast<double>(B) * Ho * Wo;
  // algorithmic bytes: dCol read (2*taps*C per px) + x read + offsets/mask read + dX written (fp32) + dOffset/dMask
  const double bytes = px * 2.0 * taps * C + static_cast<double>(B) * H * W * (2.0 * C + (dx ? 4.0 * C : 0.0)) +
                       px * 4.0 * taps * (mask ? 3 : 2) * 2.0;
  const int th = timing_begin(TC_COL2IM, bytes, static_cast<cudaStream_t>(stream));
  const size_t win_bytes = sizeof(float) * WIN_H * WIN_W * C;
  if (dx && C <= 256) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(dcn_col2im_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(sizeof(float) * WIN_H * WIN_W * 256));
      if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(dcn_col2im): %s", cudaGetErrorString(e));
      attr_done = true;
    }
    dcn_col2im_kernel<true><<<grid, GATHER_THREADS, win_bytes, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(gcol), static_cast<const __nv_bfloat16*>(x), offset, mask, dx, doffset,
        dmask, g, lddx, lddo, lddm);
  } else {
    dcn_col2im_kernel<false><<<grid, GATHER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(gcol), static_cast<const __nv_bfloat16*>(x), offset, mask, dx, doffset,
        dmask, g, lddx, lddo, lddm);
  }
  timing_end(th, static_cast<cudaStream_t>(stream));
  return check_launch("dcn_col2im");
}
