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
#include "dcn_common.cuh"
#include <stdlib.h>
#include <string.h>

#include "lsnet_internal.h"

namespace lsn {

constexpr int PATCH_H = 4, PATCH_W = 8;   // 32 output pixels per CTA: keeps the sampled rows L1-resident
constexpr int GATHER_THREADS = 256;

// grid: (patches_w, patches_h, B).  One warp per output pixel: lane k (< taps) derives the sampling position, the four
// (mask-folded) bilinear weights and the four corner pixel indices of tap k ONCE; the tap loop broadcasts them with
// shuffles, and every lane gathers/interpolates/stores the 8 channels it owns (16-byte accesses, coalesced per corner).
// Corner indices are clamped into the map (always loadable, weight 0 when outside), so the tap loop is branch-free;
// the interpolation runs on packed fp32x2 FMAs (sm_100 FFMA2).
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
  const int ldxb = static_cast<int>(g.ldx) * 2;      // pixel pitch in bytes (checked < 2^31 on the host)
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
          if (mask) m = load_mask(g, mask + p * g.ldm + grp * taps + k);
          cw0 = cn.w[0] * m; cw1 = cn.w[1] * m; cw2 = cn.w[2] * m; cw3 = cn.w[3] * m;
          co0 = static_cast<int>(cn.o[0] / g.ldx); co1 = static_cast<int>(cn.o[1] / g.ldx);
          co2 = static_cast<int>(cn.o[2] / g.ldx); co3 = static_cast<int>(cn.o[3] / g.ldx);
        }
        const int nt = min(32, taps - t0);
        for (int cb = grp * cpg; cb < (grp + 1) * cpg; cb += 256) {     // warp-uniform trip count (shuffles inside)
          const bool act = cb + lane * 8 < (grp + 1) * cpg;
          const int c0 = act ? cb + lane * 8 : cb;
          const char* xc = reinterpret_cast<const char*>(x + c0);
          for (int kk = 0; kk < nt; kk += U) {
            uint4 ld[U][4];
            float wq[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int src = min(kk + u, nt - 1);
              wq[u][0] = __shfl_sync(0xffffffffu, cw0, src); wq[u][1] = __shfl_sync(0xffffffffu, cw1, src);
              wq[u][2] = __shfl_sync(0xffffffffu, cw2, src); wq[u][3] = __shfl_sync(0xffffffffu, cw3, src);
              const int o0 = __shfl_sync(0xffffffffu, co0, src), o1 = __shfl_sync(0xffffffffu, co1, src);
              const int o2 = __shfl_sync(0xffffffffu, co2, src), o3 = __shfl_sync(0xffffffffu, co3, src);
              ld[u][0] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(o0) * ldxb));
              ld[u][1] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(o1) * ldxb));
              ld[u][2] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(o2) * ldxb));
              ld[u][3] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(o3) * ldxb));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (kk + u >= nt) continue;
              uint32_t outw[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t a0 = reinterpret_cast<const uint32_t*>(&ld[u][0])[i];
                const uint32_t a1 = reinterpret_cast<const uint32_t*>(&ld[u][1])[i];
                const uint32_t a2 = reinterpret_cast<const uint32_t*>(&ld[u][2])[i];
                const uint32_t a3 = reinterpret_cast<const uint32_t*>(&ld[u][3])[i];
                // same association as the scalar chain: ((w0*f0 + w1*f1) + w2*f2) + w3*f3, each step one fused FMA
                float2 acc = __fmul2_rn(make_float2(wq[u][0], wq[u][0]), bf16x2_f2(a0));
                acc = __ffma2_rn(make_float2(wq[u][1], wq[u][1]), bf16x2_f2(a1), acc);
                acc = __ffma2_rn(make_float2(wq[u][2], wq[u][2]), bf16x2_f2(a2), acc);
                acc = __ffma2_rn(make_float2(wq[u][3], wq[u][3]), bf16x2_f2(a3), acc);
                outw[i] = pack_bf16x2(acc.x, acc.y);
              }
              if (act)
                st_stream(dst + static_cast<long long>(t0 + kk + u) * g.C + c0, make_uint4(outw[0], outw[1], outw[2], outw[3]));
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
          if (mask) m_ = load_mask(g, mask + p * g.ldm + grp * taps + k);
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
          if (lane == kk) { my_gh = gh; my_gw = gw; my_gm = g.mask_logits ? gm * m * (1.f - m) : gm; }
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


// ---------------------------------------------------------------------------------------------------------------------
// Binned adjoint ("transposed gather") — the training default for deformable_groups == 1.
//
// The direct scatter above issues 4 vector REDs per (pixel, tap) and lane; REDG costs ~1.3 LSU cycles per lane, so it
// is bound at ~1500 cycles per output pixel whatever the math costs.  Here a CTA owns a PH x PW patch of output pixels
// and the window of the sampled map those pixels can reach with |offset| <= R:
//   A   every (pixel, tap) derives its sampling geometry once (shared memory) and counts its non-zero corners per
//       window cell; an exclusive scan turns the counts into a CSR of (column-row, weight) entries per cell;
//   C   dOffset / dMask: one warp per pixel; per tap 4 dot products D_q = <dCol, x_q> over the channels, combined
//       with the bilinear coefficients AFTER a 6-shuffle butterfly (get_coordinate_weight, ...kernel.cu:145-188);
//   B   dX: one warp per window cell walks the cell's entries, accumulates its 8 channels per lane in fp32 registers
//       and issues ONE vector RED per cell (cells overlap between neighbouring CTAs) instead of one per corner;
//   F   corners outside the window (large offsets) fall back to the direct RED.
// Same arithmetic as the reference col2im / col2im_coord kernels, different summation order.
struct BinCfg {
  int PH, PW, WH, WW, R, skip;
};

constexpr int BIN_THREADS = 256;
constexpr int BIN_WARPS = BIN_THREADS / 32;
constexpr int BIN_PT_BYTES = 5 * 16 + 4 * 8;   // shared memory per (pixel, tap): 5 records + 4 CSR entries

// Per (pixel, tap) records in shared memory (all warp-uniform when read, i.e. broadcast LDS.128):
//   ci  int4   pixel indices of the 4 corners, clamped into the map (always loadable; invalid corners get weight 0)
//   ah  float4 d(sample)/dh coefficients per corner  = mask * (-hw, -lw, +hw, +lw), 0 for invalid corners / outside
//   aw  float4 d(sample)/dw coefficients per corner  = mask * (-hh, +hh, -lh, +lh)
//   am  float4 bilinear weights per corner (dMask coefficients; times mask = the dX scatter weights)
//   gx  int4   {h0 << 16 | w0, validity / far bits, mask as float bits, 0}
template <bool DX_FP32, int MINB, int BU>
__global__ void __launch_bounds__(BIN_THREADS, MINB)
dcn_col2im_binned_kernel(const __nv_bfloat16* __restrict__ gcol, const __nv_bfloat16* __restrict__ x,
                         const float* __restrict__ offset, const float* __restrict__ mask, void* __restrict__ dx,
                         float* __restrict__ doffset, float* __restrict__ dmask, const DcnGeom g, long long lddx,
                         long long lddo, long long lddm, const BinCfg bc) {
  extern __shared__ uint4 bin_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int taps = 9;
  const int npix = bc.PH * bc.PW, npt = npix * taps, ncell = bc.WH * bc.WW;
  int4* ci = reinterpret_cast<int4*>(bin_smem);
  float4* ah = reinterpret_cast<float4*>(ci + npt);
  float4* aw = ah + npt;
  float4* am = aw + npt;
  int4* gx = reinterpret_cast<int4*>(am + npt);
  uint2* ent = reinterpret_cast<uint2*>(gx + npt);                   // CSR payload: {column row, weight}
  int* cnt = reinterpret_cast<int*>(ent + 4 * npt);                  // [ncell + 1] -> CSR row starts
  int* cur = cnt + ncell + 1;                                        // [ncell] fill cursors
  int* wsum = cur + ncell;                                           // [BIN_WARPS + 1]
  const int b = blockIdx.z;
  const int h_base = blockIdx.y * bc.PH, w_base = blockIdx.x * bc.PW;
  const int wh0 = static_cast<int>(floorf(__fmul_rn(static_cast<float>(h_base * g.sh - g.ph), g.scale_h))) - bc.R;
  const int ww0 = static_cast<int>(floorf(__fmul_rn(static_cast<float>(w_base * g.sw - g.pw), g.scale_w))) - bc.R;
  const bool want_dx = dx != nullptr;

  for (int i = tid; i <= ncell; i += BIN_THREADS) cnt[i] = 0;
  __syncthreads();
  // ---- A: geometry + per-cell counts ----
  for (int e = tid; e < npt; e += BIN_THREADS) {
    const int pix = e / taps, tap = e - pix * taps;
    const int ho = h_base + pix / bc.PW, wo = w_base + pix % bc.PW;
    float4 rh = make_float4(0.f, 0.f, 0.f, 0.f), rw = rh, rm = rh;
    int4 rc = make_int4(0, 0, 0, 0);
    float m = 1.f;
    int bits = 0, hw = 0;
    if (ho < g.Ho && wo < g.Wo) {
      const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
      float h, w;
      sample_pos(g, offset + p * g.ldo, 0, tap, ho, wo, &h, &w);
      bits = 32;
      if ((h > -1.f) && (w > -1.f) && (h < static_cast<float>(g.H)) && (w < static_cast<float>(g.W))) {
        const float hf = floorf(h), wf = floorf(w);
        const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
        const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw_ = 1.f - lw;
        if (mask) m = load_mask(g, mask + p * g.ldm + tap);
        const bool v0 = h0 >= 0 && w0 >= 0, v1 = h0 >= 0 && w0 + 1 <= g.W - 1;
        const bool v2 = h0 + 1 <= g.H - 1 && w0 >= 0, v3 = h0 + 1 <= g.H - 1 && w0 + 1 <= g.W - 1;
        bits |= 16 | (v0 ? 1 : 0) | (v1 ? 2 : 0) | (v2 ? 4 : 0) | (v3 ? 8 : 0);
        hw = (h0 << 16) | (w0 & 0xffff);
        const int ch0 = max(h0, 0), ch1 = min(h0 + 1, g.H - 1), cw0 = max(w0, 0), cw1 = min(w0 + 1, g.W - 1);
        const int r0 = (b * g.H + ch0) * g.W, r1 = (b * g.H + ch1) * g.W;
        rc = make_int4(r0 + cw0, r0 + cw1, r1 + cw0, r1 + cw1);
        // get_coordinate_weight (...kernel.cu:145-188): out-of-range corners dropped
        rh = make_float4(v0 ? -m * hw_ : 0.f, v1 ? -m * lw : 0.f, v2 ? m * hw_ : 0.f, v3 ? m * lw : 0.f);
        rw = make_float4(v0 ? -m * hh : 0.f, v1 ? m * hh : 0.f, v2 ? -m * lh : 0.f, v3 ? m * lh : 0.f);
        rm = make_float4(v0 ? hh * hw_ : 0.f, v1 ? hh * lw : 0.f, v2 ? lh * hw_ : 0.f, v3 ? lh * lw : 0.f);
        if (want_dx) {
          const float wq[4] = {rm.x, rm.y, rm.z, rm.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (wq[q] * m == 0.f) continue;
            const int r = h0 + (q >> 1) - wh0, c = w0 + (q & 1) - ww0;
            if (r >= 0 && r < bc.WH && c >= 0 && c < bc.WW) atomicAdd(&cnt[r * bc.WW + c], 1);
            else bits |= 256 << q;
          }
        }
      }
    }
    ci[e] = rc; ah[e] = rh; aw[e] = rw; am[e] = rm;
    gx[e] = make_int4(hw, bits, __float_as_int(m), 0);
  }
  __syncthreads();
  if (want_dx) {
    // ---- exclusive scan of the counts (block-wide) ----
    const int per = (ncell + BIN_THREADS - 1) / BIN_THREADS;
    const int lo = min(tid * per, ncell), hi = min(lo + per, ncell);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = lane < BIN_WARPS ? wsum[lane] : 0, iv = v;
#pragma unroll
      for (int o = 1; o < BIN_WARPS; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += t;
      }
      if (lane < BIN_WARPS) wsum[lane] = iv - v;
      if (lane == BIN_WARPS - 1) wsum[BIN_WARPS] = iv;
    }
    __syncthreads();
    int base = wsum[warp] + incl - s;
    for (int i = lo; i < hi; ++i) {
      const int c = cnt[i];
      cnt[i] = base; cur[i] = base;
      base += c;
    }
    if (tid == 0) cnt[ncell] = wsum[BIN_WARPS];
    __syncthreads();
    // ---- A2: fill the CSR ----
    for (int e = tid; e < npt; e += BIN_THREADS) {
      const int4 gg = gx[e];
      const int bits = gg.y;
      if (!(bits & 16)) continue;
      const int pix = e / taps, tap = e - pix * taps;
      const int ho = h_base + pix / bc.PW, wo = w_base + pix % bc.PW;
      const int colrow = ((b * g.Ho + ho) * g.Wo + wo) * taps + tap;
      const int h0 = gg.x >> 16, w0 = static_cast<int>(static_cast<short>(gg.x & 0xffff));
      const float4 wm = am[e];
      const float m = __int_as_float(gg.z);
      const float wq[4] = {wm.x, wm.y, wm.z, wm.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sc = wq[q] * m;
        if (((bits >> (8 + q)) & 1) || sc == 0.f) continue;
        const int r = h0 + (q >> 1) - wh0, c = w0 + (q & 1) - ww0;
        const int slot = atomicAdd(&cur[r * bc.WW + c], 1);
        ent[slot] = make_uint2(static_cast<uint32_t>(colrow), __float_as_uint(sc));
      }
    }
    __syncthreads();
  }

  // ---- C: dOffset / dMask, one warp per output pixel ----
  // Branch-free tap loop: 4 dot products <dCol, x_corner> per tap (packed fp32x2 FMAs), then the three linear
  // combinations with the warp-uniform coefficient records.  The 27 per-lane partials (gh, gw per tap, then gm per tap)
  // are reduced together by ONE transposed butterfly (31 shuffles), which leaves value i in lane i == its position in
  // the pixel's dOffset / dMask rows (coalesced stores).
  for (int pix = warp; pix < npix && !(bc.skip & 1); pix += BIN_WARPS) {
    const int ho = h_base + pix / bc.PW, wo = w_base + pix % bc.PW;
    if (ho >= g.Ho || wo >= g.Wo) continue;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    for (int cb = 0; cb < g.C; cb += 256) {
      const int c0 = cb + lane * 8;
      const bool act = c0 < g.C;
      const __nv_bfloat16* src = gcol + p * g.ldcol + (act ? c0 : 0);
      const char* xc = reinterpret_cast<const char*>(x + (act ? c0 : 0));
      const int ldxb = static_cast<int>(g.ldx) * 2;     // pixel pitch in bytes (checked < 2^31 on the host)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int4 cq = ci[pix * 9 + tap];
        const uint4 gc = ld_stream(src + tap * g.C);
        const uint4 x0 = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(cq.x) * ldxb));
        const uint4 x1 = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(cq.y) * ldxb));
        const uint4 x2 = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(cq.z) * ldxb));
        const uint4 x3 = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(cq.w) * ldxb));
        // bf16 x bf16 products accumulated in fp32 by FHFMA.BF16 (dcn_common.cuh::dot8_bf16): no operand unpacking
        const float D0 = act ? dot8_bf16(gc, x0) : 0.f, D1 = act ? dot8_bf16(gc, x1) : 0.f;
        const float D2 = act ? dot8_bf16(gc, x2) : 0.f, D3 = act ? dot8_bf16(gc, x3) : 0.f;
        const float4 kh = ah[pix * 9 + tap], kw = aw[pix * 9 + tap], km = am[pix * 9 + tap];
        v[2 * tap] += kh.x * D0 + kh.y * D1 + kh.z * D2 + kh.w * D3;
        v[2 * tap + 1] += kw.x * D0 + kw.y * D1 + kw.z * D2 + kw.w * D3;
        v[18 + tap] += km.x * D0 + km.y * D1 + km.z * D2 + km.w * D3;
      }
    }
#pragma unroll
    for (int o = 16, n = 32; o >= 1; o >>= 1, n >>= 1) {
      const bool up = lane & o;
#pragma unroll
      for (int i = 0; i < n / 2; ++i) {
        const float send = up ? v[i] : v[i + n / 2];
        const float keep = up ? v[i + n / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
    if (lane < 18) doffset[p * lddo + lane] = v[0];
    else if (lane < 27 && dmask) {
      float r = v[0];
      if (g.mask_logits) {
        const float m = __int_as_float(gx[pix * 9 + lane - 18].z);
        r *= m * (1.f - m);
      }
      dmask[p * lddm + lane - 18] = r;
    }
  }
  if (!want_dx) return;

  // ---- B: dX, one warp per window cell ----
  const float inv_ww = 1.f / static_cast<float>(bc.WW);
  for (int cell = warp; cell < ncell && !(bc.skip & 2); cell += BIN_WARPS) {
    const int s = cnt[cell], n = cnt[cell + 1] - s;
    if (n == 0) continue;
    const int r = __float2int_rz((static_cast<float>(cell) + 0.5f) * inv_ww), c = cell - r * bc.WW;
    const long long q = static_cast<long long>((b * g.H + (wh0 + r)) * g.W + (ww0 + c));
    for (int cb = 0; cb < g.C; cb += 256) {
      const int c0 = cb + lane * 8;
      const bool act = c0 < g.C;
      const __nv_bfloat16* src = gcol + (act ? c0 : 0);
      float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      int i = 0;
      for (; i + BU <= n; i += BU) {
        uint2 e[BU];
        uint4 v[BU];
#pragma unroll
        for (int u = 0; u < BU; ++u) e[u] = ent[s + i + u];
#pragma unroll
        for (int u = 0; u < BU; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(e[u].x) * g.C));
#pragma unroll
        for (int u = 0; u < BU; ++u) axpy8(__uint_as_float(e[u].y), v[u], acc);
      }
      for (; i < n; ++i) {
        const uint2 e = ent[s + i];
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(e.x) * g.C));
        axpy8(__uint_as_float(e.y), v, acc);
      }
      if (act) {
        const float av[8] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y, acc[2].x, acc[2].y, acc[3].x, acc[3].y};
        if (DX_FP32) red_f32x8(static_cast<float*>(dx) + q * lddx + c0, av);
        else red_bf16x8(static_cast<__nv_bfloat16*>(dx) + q * lddx + c0, av);
      }
    }
  }

  // ---- F: corners outside the window ----
  for (int e = warp; e < npt; e += BIN_WARPS) {
    const int4 gg = gx[e];
    const int far = (gg.y >> 8) & 15;
    if (!far) continue;
    const int pix = e / taps, tap = e - pix * taps;
    const int ho = h_base + pix / bc.PW, wo = w_base + pix % bc.PW;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    const int4 cq = ci[e];
    const int qi[4] = {cq.x, cq.y, cq.z, cq.w};
    const float4 wm = am[e];
    const float wq[4] = {wm.x, wm.y, wm.z, wm.w};
    const float m = __int_as_float(gg.z);
    const __nv_bfloat16* src = gcol + p * g.ldcol + static_cast<long long>(tap) * g.C;
    for (int c0 = lane * 8; c0 < g.C; c0 += 256) {
      float gc[8];
      bf16x8_to_float(__ldg(reinterpret_cast<const uint4*>(src + c0)), gc);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!((far >> q) & 1)) continue;
        const float sc = wq[q] * m;
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = sc * gc[t];
        if (DX_FP32) red_f32x8(static_cast<float*>(dx) + static_cast<long long>(qi[q]) * lddx + c0, v);
        else red_bf16x8(static_cast<__nv_bfloat16*>(dx) + static_cast<long long>(qi[q]) * lddx + c0, v);
      }
    }
  }
}

static int check_geom(const char* who, int C, int dg, long long ldx, long long ldcol) {
  if (dg < 1 || C % dg || (C / dg) % 8 || (ldx % 8) || (ldcol % 8))
    return set_error("%s: need C/deformable_groups %% 8 == 0 and 16-byte aligned pitches (C=%d dg=%d)", who, C, dg);
  return 0;
}

// Patch / window selection for the binned adjoint.  Returns false when the direct scatter must be used
// (deformable groups, column pitch != taps*C, > 2^31 column rows, LSNET_COL2IM=direct, or a window that does not fit).
static bool pick_binned(const DcnGeom& g, bool want_dx, BinCfg* bc, size_t* smem) {
  static int mode = -1, max_smem = 0;
  if (mode < 0) {
    const char* e = getenv("LSNET_COL2IM");
    mode = (e && !strcmp(e, "direct")) ? 0 : 1;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
#define LSN_BIN_ATTR(MB, UN)                                                                                          \
  cudaFuncSetAttribute(dcn_col2im_binned_kernel<true, MB, UN>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem); \
  cudaFuncSetAttribute(dcn_col2im_binned_kernel<false, MB, UN>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)
    LSN_BIN_ATTR(2, 8); LSN_BIN_ATTR(3, 8); LSN_BIN_ATTR(3, 4); LSN_BIN_ATTR(4, 4); LSN_BIN_ATTR(4, 8); LSN_BIN_ATTR(5, 4);
  }
  const int taps = g.kh * g.kw;
  if (!mode || g.dg != 1 || taps != 9 || g.ldcol != static_cast<long long>(taps) * g.C || g.scale_h <= 0.f || g.scale_w <= 0.f ||
      static_cast<double>(g.B) * g.Ho * g.Wo * taps >= 2147483647.0 || g.H > 32767 || g.W > 32767 ||
      g.ldx * 2 > 2147483647LL || static_cast<double>(g.B) * g.H * g.W >= 2147483647.0)
    return false;
  // Patch candidates, largest first.  4x8 measured fastest on the large levels (8x16 / 8x8: LSNET_BIN_PATCH=0/1); the
  // small pyramid levels take smaller patches so that the grid still covers the 148 SMs (their cost is the latency of
  // one CTA's serial phases, which scales with the patch size).
  static const int cand[6][2] = {{8, 16}, {8, 8}, {4, 8}, {2, 8}, {2, 4}, {1, 4}};
  const int R = 2;
  static int first = -1;
  if (first < 0) { const char* e = getenv("LSNET_BIN_PATCH"); first = e ? atoi(e) : 2; }
  for (int i = first; i < 6; ++i) {
    const int PH = cand[i][0], PW = cand[i][1];
    const long long ctas = static_cast<long long>((g.Ho + PH - 1) / PH) * ((g.Wo + PW - 1) / PW) * g.B;
    if (i < 5 && ctas < 148) continue;
    const int WH = static_cast<int>(floorf(((PH - 1) * g.sh + (g.kh - 1) * g.dh) * g.scale_h)) + 3 + 2 * R;
    const int WW = static_cast<int>(floorf(((PW - 1) * g.sw + (g.kw - 1) * g.dw) * g.scale_w)) + 3 + 2 * R;
    const long long ncell = want_dx ? static_cast<long long>(WH) * WW : 1;
    const long long npt = static_cast<long long>(PH) * PW * taps;
    const long long bytes = npt * BIN_PT_BYTES + (2 * ncell + 1 + BIN_WARPS + 1) * 4 + 16;
    if (bytes > 72 * 1024 || bytes > max_smem) continue;   // <= 72 KB keeps 3 CTAs per SM
    static int skip = -1;
    if (skip < 0) { const char* e = getenv("LSNET_BIN_SKIP"); skip = e ? atoi(e) : 0; }
    *bc = BinCfg{PH, PW, want_dx ? WH : 1, want_dx ? WW : 1, R, skip};
    *smem = static_cast<size_t>(bytes);
    return true;
  }
  return false;
}

}  // namespace lsn

using namespace lsn;

extern "C" int lsnet_dcn_im2col_bf16(const void* x, int B, int H, int W, int C, long long ldx, const float* offset,
                                     long long ldo, const float* mask, long long ldm, int Ho, int Wo, int kh, int kw,
                                     int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                                     float scale_h, float scale_w, int deformable_groups, void* col, long long ldcol,
                                     int mask_logits, void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if (int rc = check_geom("lsnet_dcn_im2col_bf16", C, deformable_groups, ldx, ldcol)) return rc;
  DcnGeom g{B, H, W, C, Ho, Wo, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, scale_h, scale_w,
            deformable_groups, ldx, ldo, ldm, ldcol, (mask && mask_logits) ? 1 : 0};
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

namespace lsn {
// TMA-staged adjoint (dcn_adjoint.cu): 0 = launched, 1 = shape not taken, < 0 = error
int dcn_adjoint_tma(const DcnGeom& g, const void* gcol, const void* x, const float* offset, const float* mask, void* dx,
                    long long lddx, int dx_fp32, float* doffset, long long lddo, float* dmask, long long lddm,
                    cudaStream_t st);
}

extern "C" int lsnet_dcn_col2im_bf16(const void* gcol, long long ldcol, const void* x, int B, int H, int W, int C,
                                     long long ldx, const float* offset, long long ldo, const float* mask,
                                     long long ldm, int Ho, int Wo, int kh, int kw, int stride_h, int stride_w,
                                     int pad_h, int pad_w, int dil_h, int dil_w, float scale_h, float scale_w,
                                     int deformable_groups, void* dx, long long lddx, int dx_fp32, float* doffset,
                                     long long lddo, float* dmask, long long lddm, int mask_logits, void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if (int rc = check_geom("lsnet_dcn_col2im_bf16", C, deformable_groups, ldx, ldcol)) return rc;
  if (dx && (lddx % 8)) return set_error("lsnet_dcn_col2im_bf16: dx pitch must be a multiple of 8 bf16");
  DcnGeom g{B, H, W, C, Ho, Wo, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, scale_h, scale_w,
            deformable_groups, ldx, ldo, ldm, ldcol, (mask && mask_logits) ? 1 : 0};
  dim3 grid((Wo + PATCH_W - 1) / PATCH_W, (Ho + PATCH_H - 1) / PATCH_H, B);
  const double taps = kh * kw, px = static_cast<double>(B) * Ho * Wo;
  // algorithmic bytes: dCol read (2*taps*C per px) + x read + offsets/mask read + dX written (bf16) + dOffset/dMask
  const double bytes = px * 2.0 * taps * C + static_cast<double>(B) * H * W * (2.0 * C + (dx ? (dx_fp32 ? 4.0 : 2.0) * C : 0.0)) +
                       px * 4.0 * taps * (mask ? 3 : 2) * 2.0;
  const int th = timing_begin(TC_COL2IM, bytes, static_cast<cudaStream_t>(stream));
  BinCfg bc;
  size_t bin_smem = 0;
  const int rc_tma = dcn_adjoint_tma(g, gcol, x, offset, mask, dx, lddx, dx_fp32, doffset, lddo, dmask, lddm,
                                     static_cast<cudaStream_t>(stream));
  if (rc_tma < 0) return 1;
  if (rc_tma == 0) {
    // dCol slices staged by TMA (dcn_adjoint.cu)
  } else if (pick_binned(g, dx != nullptr, &bc, &bin_smem)) {
    dim3 bgrid((Wo + bc.PW - 1) / bc.PW, (Ho + bc.PH - 1) / bc.PH, B);
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("LSNET_BIN_VARIANT"); variant = e ? atoi(e) : 44; }
#define LSN_BIN_LAUNCH(F32, MB, UN)                                                                                   \
  dcn_col2im_binned_kernel<F32, MB, UN><<<bgrid, BIN_THREADS, bin_smem, static_cast<cudaStream_t>(stream)>>>(        \
      static_cast<const __nv_bfloat16*>(gcol), static_cast<const __nv_bfloat16*>(x), offset, mask, dx, doffset, dmask, g, \
      lddx, lddo, lddm, bc)
#define LSN_BIN_VARIANT(MB, UN)                                        \
  if (variant == MB * 10 + UN) {                                       \
    if (dx_fp32) LSN_BIN_LAUNCH(true, MB, UN); else LSN_BIN_LAUNCH(false, MB, UN); \
  }
    LSN_BIN_VARIANT(2, 8) else LSN_BIN_VARIANT(3, 8) else LSN_BIN_VARIANT(3, 4) else LSN_BIN_VARIANT(4, 4)
    else LSN_BIN_VARIANT(4, 8) else LSN_BIN_VARIANT(5, 4) else return set_error("lsnet_dcn_col2im_bf16: unknown LSNET_BIN_VARIANT");
  } else if (dx_fp32)
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
