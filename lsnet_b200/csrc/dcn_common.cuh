// Device helpers shared by the deformable-convolution kernels (dcn_gather.cu, dcn_fused.cu): sampling geometry of
// DCNv1 / DCNv2 / pyramid DCN (reference arithmetic: mmdet/ops/dcn/src/cuda/deform_conv_cuda_kernel.cu:108-188, 245-297,
// 847-910) and the bf16 <-> fp32x2 operand forms of the packed FMA interpolation.
#pragma once
#include "common.cuh"

namespace lsn {

struct DcnGeom {
  int B, H, W, C;          // input (sampled) feature map, NHWC
  int Ho, Wo;              // sampling / output grid
  int kh, kw, sh, sw, ph, pw, dh, dw;
  float scale_h, scale_w;  // pyramid: base grid is scaled, the learned offset is not (…kernel.cu:281-282)
  int dg;                  // deformable groups
  long long ldx, ldo, ldm, ldcol;
  int mask_logits;         // mask holds the raw conv_offset logits: m = sigmoid(raw), dMask is returned w.r.t. the logits
};

struct Corner {
  float w[4];       // bilinear weights (0 when the corner is outside)
  long long o[4];   // pixel offsets (elements / ldx) of the 4 corners (clamped in range)
  bool inside;
  float lh, lw;
  bool v[4];
};

__device__ __forceinline__ float load_mask(const DcnGeom& g, const float* mp) {
  const float r = __ldg(mp);
  return g.mask_logits ? 1.f / (1.f + __expf(-r)) : r;
}

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

// Bilinear corners of a sampling position as PIXEL indices (b*H + h)*W + w (host guarantees B*H*W < 2^31), clamped into
// the map so that they are always loadable, with weight 0 for corners outside (same arithmetic as make_corner).
__device__ __forceinline__ void corner_pixels(const DcnGeom& g, int b, float h, float w, uint32_t (&pi)[4], float (&wt)[4]) {
  const bool inside = (h > -1.f) && (w > -1.f) && (h < static_cast<float>(g.H)) && (w < static_cast<float>(g.W));
  const float hf = floorf(h), wf = floorf(w);
  const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
  const int h1 = h0 + 1, w1 = w0 + 1;
  const float lh = h - hf, lw = w - wf, hh = 1.f - lh, hw = 1.f - lw;
  const bool t = inside && h0 >= 0, bt = inside && h1 <= g.H - 1, l = w0 >= 0, r = w1 <= g.W - 1;
  wt[0] = (t && l) ? hh * hw : 0.f;
  wt[1] = (t && r) ? hh * lw : 0.f;
  wt[2] = (bt && l) ? lh * hw : 0.f;
  wt[3] = (bt && r) ? lh * lw : 0.f;
  const int ch0 = min(max(h0, 0), g.H - 1), ch1 = min(max(h1, 0), g.H - 1);
  const int cw0 = min(max(w0, 0), g.W - 1), cw1 = min(max(w1, 0), g.W - 1);
  const int r0 = (b * g.H + ch0) * g.W, r1 = (b * g.H + ch1) * g.W;
  pi[0] = static_cast<uint32_t>(r0 + cw0);
  pi[1] = static_cast<uint32_t>(r0 + cw1);
  pi[2] = static_cast<uint32_t>(r1 + cw0);
  pi[3] = static_cast<uint32_t>(r1 + cw1);
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

// bf16x2 word -> (lo, hi) as a float2 register pair (one shift, one mask), the operand form of the packed fp32x2 FMA
__device__ __forceinline__ float2 bf16x2_f2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// <a, b> over 8 bf16 channels on the mixed-precision FMA of sm_100 (PTX `fma.rn.f32.bf16` -> SASS `FHFMA.BF16 Rd, Ra.H0|H1,
// Rb.H0|H1, Rc`: fp32 accumulator, both multiplicands taken straight from the halves of packed bf16x2 registers).  No
// unpack instructions at all: 8 FHFMA in two independent chains instead of 8 + 8 shifts / masks and 4 FFMA2.  The
// products are exact in fp32 either way, so the result differs from dot8 only by summation order.
__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  float lo = 0.f, hi = 0.f;
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %2;\n\tmov.b32 {bl, bh}, %3;\n\t"
        "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %1, ah, bh, %1;\n\t}"
        : "+f"(lo), "+f"(hi)
        : "r"(pa[i]), "r"(pb[i]));
  return lo + hi;
}

__device__ __forceinline__ void axpy8(float w, const uint4& a, float2 (&acc)[4]) {
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = __ffma2_rn(w2, bf16x2_f2(pa[i]), acc[i]);
}

}  // namespace lsn
