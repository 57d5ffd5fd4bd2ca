// Adjoint of the deformable-convolution sampling, TMA-staged (sm_100a):  dCol [pixels, 9*C] (= dY . W, bf16) ->
//   dX      (scatter of the bilinear weights, reference: *_col2im kernels, deform_conv_cuda_kernel.cu:333-448, 912-970)
//   dOffset, dMask (channel reductions, reference: *_col2im_coord kernels, :486-615, 972-1044)
// for 3x3 taps, deformable_groups == 1, C % 64 == 0 (every LSNet head site); other shapes keep dcn_col2im*.
//
// A CTA owns a patch of <= 32 output pixels and walks the channels in 64-wide blocks.  Per block ONE 5-D TMA box
// ([64 ch, 9 taps, PW, PH, 1] of dCol viewed as [B, Ho, Wo, 9, C], SWIZZLE_128B, mbarrier completion; three CTAs per SM
// cover each other's load latency) stages the slice dCol[patch, 9 taps, 64 ch] in shared memory -- every later read of dCol (once for the reductions,
// once per bilinear corner for dX) is a conflict-free 16-byte shared-memory read instead of an L2 round trip, and dCol
// crosses HBM/L2 exactly once.  Eight lanes own a 128-byte row (lane = 8 channels):
//   A  (once)      sampling geometry of every (pixel, tap): corner pixel indices, (lh, lw, mask, validity) records, and a
//                  CSR of (slice row, weight) entries per cell of the patch's window of the sampled map (|offset| <= R)
//   C  (per block) lane group g = pixel g: 4 dot products <dCol, x_corner> per tap, software pipelined (the corner loads
//                  of tap t+1 fly while tap t is reduced); the 8 lanes' partials are folded with 4 shuffles per tap and
//                  accumulated over the blocks in a small shared-memory table (keeps the kernel at 80 registers)
//   B  (per block) lane groups pull window cells from a shared counter (cells differ in their number of entries): walk
//                  the cell's CSR entries, accumulate in fp32 registers, ONE vector RED per cell and block
//   F  (per block) corners outside the window (large offsets): direct REDs
//   end            combination of the 4 sums per (pixel, tap) with the coefficient records (get_coordinate_weight,
//                  ...kernel.cu:145-188) -> dOffset / dMask rows
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "dcn_common.cuh"
#include "lsnet_internal.h"

namespace lsn {

#ifndef LSN_FHFMA
#define LSN_FHFMA 1      // dOffset / dMask dot products on FHFMA.BF16 (0: unpack + FFMA2, the A/B build)
#endif

constexpr int ADJ_THREADS = 256;
constexpr int ADJ_GROUPS = ADJ_THREADS / 8;     // 32 lane groups of 8
constexpr int ADJ_TAPS = 9;

struct AdjCfg {
  int PH, PW, pw_shift, WH, WW, R;
};

__device__ __forceinline__ float2 dot_acc(const float2 (&a)[4], const uint4& b) {
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
  float2 acc = __fmul2_rn(a[0], bf16x2_f2(pb[0]));
#pragma unroll
  for (int i = 1; i < 4; ++i) acc = __ffma2_rn(a[i], bf16x2_f2(pb[i]), acc);
  return acc;
}

__device__ __forceinline__ void red_add_bf16x8(__nv_bfloat16* dst, const float2 (&v)[4]) {
  const uint32_t a = pack_bf16x2(v[0].x, v[0].y), b = pack_bf16x2(v[1].x, v[1].y), c = pack_bf16x2(v[2].x, v[2].y),
                 d = pack_bf16x2(v[3].x, v[3].y);
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void red_add_f32x8(float* dst, const float2 (&v)[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y)
               : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x),
               "f"(v[3].y)
               : "memory");
}

template <bool DX_FP32>
__global__ void __launch_bounds__(ADJ_THREADS, 3)
dcn_adjoint_tma_kernel(const __grid_constant__ CUtensorMap tmCol, const __nv_bfloat16* __restrict__ x,
                       const float* __restrict__ offset, const float* __restrict__ mask, void* __restrict__ dx,
                       float* __restrict__ doffset, float* __restrict__ dmask, const DcnGeom g, long long lddx,
                       long long lddo, long long lddm, const AdjCfg bc) {
  extern __shared__ uint8_t adj_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(adj_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = tid >> 3, sub = tid & 7;
  const int npix = bc.PH * bc.PW, npt = npix * ADJ_TAPS, ncell = bc.WH * bc.WW;
  const int slice_bytes = npt * 128;
  uint4* ci = reinterpret_cast<uint4*>(smem + slice_bytes);           // corner pixel indices per (pixel, tap)
  float4* cf = reinterpret_cast<float4*>(ci + npt);                   // {lh, lw, mask, bits}
  float4* dsum = cf + npt;                                            // <dCol, x_corner> sums per (pixel, tap)
  int* hwv = reinterpret_cast<int*>(dsum + npt);                      // h0 << 16 | w0
  uint2* ent = reinterpret_cast<uint2*>(hwv + npt + (npt & 1));       // CSR payload: {swizzled slice row offset, weight}
  int* cnt = reinterpret_cast<int*>(ent + 4 * npt);                   // [ncell + 1] CSR row starts
  int* cur = cnt + ncell + 1;                                         // [ncell] fill cursors
  int* wsum = cur + ncell;                                            // [9]
  int* next_cell = wsum + 9;                                          // [2] work counters of phase B (alternating)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(next_cell + 2) + 7) & ~uintptr_t(7));
  const uint32_t slice_s = smem_u32(smem), ci_s = smem_u32(ci), cf_s = smem_u32(cf), ent_s = smem_u32(ent),
                 dsum_s = smem_u32(dsum);

  const int b = blockIdx.z;
  const int h_base = blockIdx.y * bc.PH, w_base = blockIdx.x * bc.PW;
  const int wh0 = static_cast<int>(floorf(__fmul_rn(static_cast<float>(h_base * g.sh - g.ph), g.scale_h))) - bc.R;
  const int ww0 = static_cast<int>(floorf(__fmul_rn(static_cast<float>(w_base * g.sw - g.pw), g.scale_w))) - bc.R;
  const bool want_dx = dx != nullptr;
  const int ncb = g.C / 64;

  if (tid == 0) {
    tma_prefetch_desc(&tmCol);
    mbar_init(full_bar, 1);
    fence_barrier_init();
    next_cell[0] = ADJ_GROUPS;
    next_cell[1] = ADJ_GROUPS;
  }
  for (int i = tid; i <= ncell; i += ADJ_THREADS) cnt[i] = 0;
  __syncthreads();
  if (tid == 0) {      // the first slice flies while the geometry is derived
    mbar_expect_tx(full_bar, slice_bytes);
    tma_load_5d(smem, &tmCol, full_bar, 0, 0, w_base, h_base, b);
  }

  // ---- A: geometry + per-cell counts ----
  for (int e = tid; e < npt; e += ADJ_THREADS) {
    const int pix = e / ADJ_TAPS, tap = e - pix * ADJ_TAPS;
    const int ho = h_base + (pix >> bc.pw_shift), wo = w_base + (pix & (bc.PW - 1));
    uint4 rc = make_uint4(0u, 0u, 0u, 0u);
    float lh = 0.f, lw = 0.f, m = 1.f;
    int bits = 0, hw = 0;
    if (ho < g.Ho && wo < g.Wo) {
      const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
      float h, w;
      sample_pos(g, offset + p * g.ldo, 0, tap, ho, wo, &h, &w);
      bits = 32;
      if ((h > -1.f) && (w > -1.f) && (h < static_cast<float>(g.H)) && (w < static_cast<float>(g.W))) {
        const float hf = floorf(h), wf = floorf(w);
        const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
        lh = h - hf; lw = w - wf;
        const float hh = 1.f - lh, hw_ = 1.f - lw;
        if (mask) m = load_mask(g, mask + p * g.ldm + tap);
        const bool v0 = h0 >= 0 && w0 >= 0, v1 = h0 >= 0 && w0 + 1 <= g.W - 1;
        const bool v2 = h0 + 1 <= g.H - 1 && w0 >= 0, v3 = h0 + 1 <= g.H - 1 && w0 + 1 <= g.W - 1;
        bits |= 16 | (v0 ? 1 : 0) | (v1 ? 2 : 0) | (v2 ? 4 : 0) | (v3 ? 8 : 0);
        hw = (h0 << 16) | (w0 & 0xffff);
        const int ch0 = max(h0, 0), ch1 = min(h0 + 1, g.H - 1), cw0 = max(w0, 0), cw1 = min(w0 + 1, g.W - 1);
        const int r0 = (b * g.H + ch0) * g.W, r1 = (b * g.H + ch1) * g.W;
        rc = make_uint4(r0 + cw0, r0 + cw1, r1 + cw0, r1 + cw1);
        if (want_dx) {
          const float wq[4] = {v0 ? hh * hw_ : 0.f, v1 ? hh * lw : 0.f, v2 ? lh * hw_ : 0.f, v3 ? lh * lw : 0.f};
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
    ci[e] = rc;
    cf[e] = make_float4(lh, lw, m, __int_as_float(bits));
    dsum[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    hwv[e] = hw;
  }
  __syncthreads();
  if (want_dx) {
    // ---- exclusive scan of the counts (block-wide), then the CSR fill ----
    const int per = (ncell + ADJ_THREADS - 1) / ADJ_THREADS;
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
      int v = lane < 8 ? wsum[lane] : 0, iv = v;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, iv, o);
        if (lane >= o) iv += t;
      }
      if (lane < 8) wsum[lane] = iv - v;
      if (lane == 7) wsum[8] = iv;
    }
    __syncthreads();
    int base = wsum[warp] + incl - s;
    for (int i = lo; i < hi; ++i) {
      const int c = cnt[i];
      cnt[i] = base; cur[i] = base;
      base += c;
    }
    if (tid == 0) cnt[ncell] = wsum[8];
    __syncthreads();
    for (int e = tid; e < npt; e += ADJ_THREADS) {
      const float4 f = cf[e];
      const int bits = __float_as_int(f.w);
      if (!(bits & 16)) continue;
      const int hw = hwv[e];
      const int h0 = hw >> 16, w0 = static_cast<int>(static_cast<short>(hw & 0xffff));
      const float hh = 1.f - f.x, hw_ = 1.f - f.y;
      const float wq[4] = {(bits & 1) ? hh * hw_ : 0.f, (bits & 2) ? hh * f.y : 0.f, (bits & 4) ? f.x * hw_ : 0.f,
                           (bits & 8) ? f.x * f.y : 0.f};
      // byte offset of slice row e with the row's swizzle term folded in: a lane's 16 bytes sit at (off ^ (sub << 4))
      const uint32_t off = static_cast<uint32_t>(e) * 128u + ((static_cast<uint32_t>(e) & 7u) << 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float sc = wq[q] * f.z;
        if (((bits >> (8 + q)) & 1) || sc == 0.f) continue;
        const int r = h0 + (q >> 1) - wh0, c = w0 + (q & 1) - ww0;
        const int slot = atomicAdd(&cur[r * bc.WW + c], 1);
        ent[slot] = make_uint2(off, __float_as_uint(sc));
      }
    }
    __syncthreads();
  }

  // ---- channel blocks ----
  const uint32_t ldxb = static_cast<uint32_t>(g.ldx) * 2u;
  const float inv_ww = 1.f / static_cast<float>(bc.WW);
  const bool c_active = gid < npix;
  const int r0 = gid * ADJ_TAPS;
  const uint32_t sub16 = static_cast<uint32_t>(sub) << 4;

  for (int cb = 0; cb < ncb; ++cb) {
    mbar_wait(full_bar, cb & 1);
    const uint32_t sl = slice_s;

    // ---- C: <dCol, x_corner> per tap, pipelined over the taps ----
    if (c_active) {
      const unsigned long long xc = reinterpret_cast<unsigned long long>(x + cb * 64 + sub * 8);
      uint4 dcA, dcB, xA[4], xB[4];
#define LSN_ADJ_LOAD(dc, xs, t)                                                       \
  {                                                                                   \
    const int r_ = r0 + (t);                                                          \
    dc = lds128(sl + r_ * 128 + ((sub ^ (r_ & 7)) << 4));                             \
    const uint4 c_ = lds128(ci_s + r_ * 16);                                          \
    xs[0] = ldg128_at(xc, c_.x, ldxb);                                                \
    xs[1] = ldg128_at(xc, c_.y, ldxb);                                                \
    xs[2] = ldg128_at(xc, c_.z, ldxb);                                                \
    xs[3] = ldg128_at(xc, c_.w, ldxb);                                                \
  }
#if LSN_FHFMA
#define LSN_ADJ_DOT4(dc, xs, d_) \
  _Pragma("unroll") for (int q_ = 0; q_ < 4; ++q_) d_[q_] = dot8_bf16(dc, xs[q_]);
#else
#define LSN_ADJ_DOT4(dc, xs, d_)                                                      \
  {                                                                                   \
    const float2 ga_[4] = {bf16x2_f2(dc.x), bf16x2_f2(dc.y), bf16x2_f2(dc.z), bf16x2_f2(dc.w)}; \
    _Pragma("unroll") for (int q_ = 0; q_ < 4; ++q_) {                                \
      const float2 t_ = dot_acc(ga_, xs[q_]);                                         \
      d_[q_] = t_.x + t_.y;                                                           \
    }                                                                                 \
  }
#endif
      // 4 dot products of 8 channels each, folded over the 8 lanes of the group with 4 shuffles (transposed butterfly:
      // lanes 0-1 end with corner 0, 2-3 with corner 1, ...), then one lane per corner adds into the shared table
#define LSN_ADJ_DOT(dc, xs, t)                                                        \
  {                                                                                   \
    float d_[4];                                                                      \
    LSN_ADJ_DOT4(dc, xs, d_)                                                          \
    const bool u4_ = sub & 4, u2_ = sub & 2;                                          \
    const float a0_ = (u4_ ? d_[2] : d_[0]) + __shfl_xor_sync(0xffffffffu, u4_ ? d_[0] : d_[2], 4); \
    const float a1_ = (u4_ ? d_[3] : d_[1]) + __shfl_xor_sync(0xffffffffu, u4_ ? d_[1] : d_[3], 4); \
    float v_ = (u2_ ? a1_ : a0_) + __shfl_xor_sync(0xffffffffu, u2_ ? a0_ : a1_, 2);  \
    v_ += __shfl_xor_sync(0xffffffffu, v_, 1);                                        \
    if (!(sub & 1)) {                                                                 \
      const uint32_t a_ = dsum_s + (r0 + (t)) * 16 + (sub >> 1) * 4;                  \
      float o_;                                                                       \
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o_) : "r"(a_));                   \
      o_ += v_;                                                                       \
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_), "f"(o_) : "memory");          \
    }                                                                                 \
  }
      LSN_ADJ_LOAD(dcA, xA, 0)
      LSN_ADJ_LOAD(dcB, xB, 1)
      LSN_ADJ_DOT(dcA, xA, 0)
      LSN_ADJ_LOAD(dcA, xA, 2)
      LSN_ADJ_DOT(dcB, xB, 1)
      LSN_ADJ_LOAD(dcB, xB, 3)
      LSN_ADJ_DOT(dcA, xA, 2)
      LSN_ADJ_LOAD(dcA, xA, 4)
      LSN_ADJ_DOT(dcB, xB, 3)
      LSN_ADJ_LOAD(dcB, xB, 5)
      LSN_ADJ_DOT(dcA, xA, 4)
      LSN_ADJ_LOAD(dcA, xA, 6)
      LSN_ADJ_DOT(dcB, xB, 5)
      LSN_ADJ_LOAD(dcB, xB, 7)
      LSN_ADJ_DOT(dcA, xA, 6)
      LSN_ADJ_LOAD(dcA, xA, 8)
      LSN_ADJ_DOT(dcB, xB, 7)
      LSN_ADJ_DOT(dcA, xA, 8)
#undef LSN_ADJ_LOAD
#undef LSN_ADJ_DOT
#undef LSN_ADJ_DOT4
    }

    if (want_dx) {
      // ---- B: dX, lane groups pull window cells from a shared counter ----
      int* ctr = &next_cell[cb & 1];
      int cell = gid;
      while (true) {
        const bool act = cell < ncell;
        if (!__any_sync(0xffffffffu, act)) break;       // warp-uniform exit (the shuffle below needs every lane)
        if (act) {
          const int s = cnt[cell], n = cnt[cell + 1] - s;
          if (n > 0) {
            float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            uint32_t ea = ent_s + s * 8;
            int i = 0;
            for (; i + 2 <= n; i += 2, ea += 16) {
              uint32_t e0r, e0w, e1r, e1w;
              asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e0r), "=r"(e0w) : "r"(ea));
              asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e1r), "=r"(e1w) : "r"(ea + 8));
              const uint4 v0 = lds128(sl + (e0r ^ sub16));
              const uint4 v1 = lds128(sl + (e1r ^ sub16));
              axpy8(__uint_as_float(e0w), v0, acc);
              axpy8(__uint_as_float(e1w), v1, acc);
            }
            if (i < n) {
              uint32_t er, ew;
              asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(er), "=r"(ew) : "r"(ea));
              const uint4 v = lds128(sl + (er ^ sub16));
              axpy8(__uint_as_float(ew), v, acc);
            }
            const int r = __float2int_rz((static_cast<float>(cell) + 0.5f) * inv_ww), c = cell - r * bc.WW;
            const long long qp = static_cast<long long>((b * g.H + (wh0 + r)) * g.W + (ww0 + c));
            if (DX_FP32) red_add_f32x8(static_cast<float*>(dx) + qp * lddx + cb * 64 + sub * 8, acc);
            else red_add_bf16x8(static_cast<__nv_bfloat16*>(dx) + qp * lddx + cb * 64 + sub * 8, acc);
          }
        }
        int nxt = ncell;
        if (act && sub == 0) nxt = atomicAdd(ctr, 1);
        cell = __shfl_sync(0xffffffffu, nxt, lane & ~7);
      }
      // ---- F: corners outside the window ----
      for (int e = gid; e < npt; e += ADJ_GROUPS) {
        const float4 f = lds128f(cf_s + e * 16);
        const int bits = __float_as_int(f.w);
        const int far = (bits >> 8) & 15;
        if (!far) continue;
        const uint4 cq = lds128(ci_s + e * 16);
        const uint32_t qi[4] = {cq.x, cq.y, cq.z, cq.w};
        const float hh = 1.f - f.x, hw_ = 1.f - f.y;
        const float wq[4] = {hh * hw_, hh * f.y, f.x * hw_, f.x * f.y};
        const uint4 v = lds128(sl + e * 128 + ((sub ^ (e & 7)) << 4));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!((far >> q) & 1)) continue;
          float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          axpy8(wq[q] * f.z, v, acc);
          if (DX_FP32) red_add_f32x8(static_cast<float*>(dx) + static_cast<long long>(qi[q]) * lddx + cb * 64 + sub * 8, acc);
          else red_add_bf16x8(static_cast<__nv_bfloat16*>(dx) + static_cast<long long>(qi[q]) * lddx + cb * 64 + sub * 8, acc);
        }
      }
    }
    __syncthreads();       // every read of the slice is done: refill it with the next block
    if (tid == 0) {
      next_cell[(cb & 1) ^ 1] = ADJ_GROUPS;     // the counter the block after next will use
      if (cb + 1 < ncb) {
        mbar_expect_tx(full_bar, slice_bytes);
        tma_load_5d(smem, &tmCol, full_bar, (cb + 1) * 64, 0, w_base, h_base, b);
      }
    }
  }

  // ---- end: combine the 4 sums of every (pixel, tap) with the coefficient records, write the rows ----
  // (the last __syncthreads above ordered the shared-memory accumulation before these reads)
  for (int e = tid; e < npt; e += ADJ_THREADS) {
    const int pix = e / ADJ_TAPS, tap = e - pix * ADJ_TAPS;
    const int ho = h_base + (pix >> bc.pw_shift), wo = w_base + (pix & (bc.PW - 1));
    if (ho >= g.Ho || wo >= g.Wo) continue;
    const long long p = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
    const float4 f = cf[e], d = dsum[e];
    const int bits = __float_as_int(f.w);
    const float lh = f.x, lw = f.y, m = f.z, hh = 1.f - lh, hw_ = 1.f - lw;
    const float v0 = (bits & 1) ? 1.f : 0.f, v1 = (bits & 2) ? 1.f : 0.f, v2 = (bits & 4) ? 1.f : 0.f,
                v3 = (bits & 8) ? 1.f : 0.f;
    // get_coordinate_weight (...kernel.cu:145-188): out-of-range corners dropped
    const float gh = m * (-hw_ * v0 * d.x - lw * v1 * d.y + hw_ * v2 * d.z + lw * v3 * d.w);
    const float gw = m * (-hh * v0 * d.x + hh * v1 * d.y - lh * v2 * d.z + lh * v3 * d.w);
    float gm = hh * hw_ * v0 * d.x + hh * lw * v1 * d.y + lh * hw_ * v2 * d.z + lh * lw * v3 * d.w;
    doffset[p * lddo + 2 * tap] = gh;
    doffset[p * lddo + 2 * tap + 1] = gw;
    if (dmask) {
      if (g.mask_logits) gm *= m * (1.f - m);
      dmask[p * lddm + tap] = gm;
    }
  }
}

// Patch / window selection.  false: the shape goes to the older kernels (dcn_gather.cu).
static bool pick_adjoint(const DcnGeom& g, bool want_dx, AdjCfg* bc, size_t* smem) {
  static int on = -1, max_smem = 0;
  if (on < 0) {
    const char* e = getenv("LSNET_ADJOINT_TMA");
    on = e ? atoi(e) : 1;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(dcn_adjoint_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(dcn_adjoint_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  }
  const int taps = g.kh * g.kw;
  // on == 1 (default): pyramid calls that sample a FINER level (scale 2: 4x the window cells per patch, most of them with
  // one or two entries) stay on the binned kernel of dcn_gather.cu -- r02: 0.144 vs 0.114 ms at 50x84 <- 100x168;
  // on == 2 forces this kernel for every supported shape.
  if (on == 1 && g.scale_h * g.scale_w > 2.25f) return false;
  if (!on || g.dg != 1 || taps != ADJ_TAPS || g.kh != 3 || g.C % 64 || g.ldcol % 8 || g.scale_h <= 0.f || g.scale_w <= 0.f ||
      static_cast<double>(g.B) * g.Ho * g.Wo * taps >= 2147483647.0 || g.H > 32767 || g.W > 32767 ||
      g.ldx * 2 > 2147483647LL || static_cast<double>(g.B) * g.H * g.W >= 2147483647.0 || g.B > 65535)
    return false;
  static const int cand[4][2] = {{4, 8}, {2, 8}, {2, 4}, {1, 4}};
  const int R = 2;
  for (int i = 0; i < 4; ++i) {
    const int PH = cand[i][0], PW = cand[i][1];
    const long long ctas = static_cast<long long>((g.Ho + PH - 1) / PH) * ((g.Wo + PW - 1) / PW) * g.B;
    if (i < 3 && ctas < 2 * num_sms()) continue;      // small levels: smaller patches so the grid still covers the SMs
    const int WH = static_cast<int>(floorf(((PH - 1) * g.sh + (g.kh - 1) * g.dh) * g.scale_h)) + 3 + 2 * R;
    const int WW = static_cast<int>(floorf(((PW - 1) * g.sw + (g.kw - 1) * g.dw) * g.scale_w)) + 3 + 2 * R;
    const long long ncell = want_dx ? static_cast<long long>(WH) * WW : 1;
    const long long npt = static_cast<long long>(PH) * PW * taps;
    const long long bytes = 1024 + npt * 128 + npt * (16 + 16 + 16 + 4 + 32) + 8 + (2 * ncell + 1 + 9 + 2) * 4 + 64;
    if (bytes > 74 * 1024 || bytes > max_smem) continue;      // <= 74 KB keeps 3 CTAs per SM
    int sh = 0;
    while ((1 << sh) < PW) ++sh;
    *bc = AdjCfg{PH, PW, sh, want_dx ? WH : 1, want_dx ? WW : 1, R};
    *smem = static_cast<size_t>(bytes);
    return true;
  }
  return false;
}

// returns 1 when the shape is not taken (caller falls back), 0 on success, < 0 on error
int dcn_adjoint_tma(const DcnGeom& g, const void* gcol, const void* x, const float* offset, const float* mask, void* dx,
                    long long lddx, int dx_fp32, float* doffset, long long lddo, float* dmask, long long lddm,
                    cudaStream_t st) {
  AdjCfg bc;
  size_t smem = 0;
  if (!pick_adjoint(g, dx != nullptr, &bc, &smem)) return 1;
  CUtensorMap tm;
  if (make_map_col5d(&tm, gcol, g.B, g.Ho, g.Wo, ADJ_TAPS, g.C, g.ldcol, bc.PW, bc.PH)) return -1;
  dim3 grid((g.Wo + bc.PW - 1) / bc.PW, (g.Ho + bc.PH - 1) / bc.PH, g.B);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  if (dx_fp32)
    dcn_adjoint_tma_kernel<true><<<grid, ADJ_THREADS, smem, st>>>(tm, xp, offset, mask, dx, doffset, dmask, g, lddx, lddo, lddm, bc);
  else
    dcn_adjoint_tma_kernel<false><<<grid, ADJ_THREADS, smem, st>>>(tm, xp, offset, mask, dx, doffset, dmask, g, lddx, lddo, lddm, bc);
  return 0;
}

}  // namespace lsn
