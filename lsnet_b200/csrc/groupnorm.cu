// GroupNorm (+ optional residual add, + optional ReLU) over pixel-major bf16 maps, forward and backward (sm_100a).
// Replaces the nn.GroupNorm calls that sit between every pair of hot kernels in FPN / LSHead
// (mmdet/models/necks/fpn.py:117-133 ConvModule norm, dense_heads/lsnet_head.py:97-113,700-708,1843) — SURVEY §8 row f1.
// HBM-bound streaming kernels: forward reads x twice (statistics, apply) and writes y once; backward reads x, dy twice
// and writes dx once.  Statistics: fp32 partial sums per CTA, fp64 atomics across CTAs (no cancellation trouble for
// E[x^2]-E[x]^2 at 134k elements per group), one thread = one 16-byte vector = 8 channels of ONE group.
#include "common.cuh"
#include "lsnet_internal.h"

namespace lsn {

constexpr int GN_THREADS = 256;
constexpr int GN_PIX_PER_CTA = 128;

struct GnArgs {
  int B, HW, C, G;          // C/G % 8 == 0
  long long ldx, ldx2, ldy;
  float eps;
  int relu;
};

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                            pack_bf16x2(f[6], f[7]));
}

// grid (ceil(HW / GN_PIX_PER_CTA), B).  stats[b][g] = {sum, sumsq} (double, pre-zeroed)
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x2, const GnArgs a,
                double* __restrict__ stats, unsigned* __restrict__ ticket, float2* __restrict__ mr, double n_elem) {
  extern __shared__ float sm[];   // [vecs_per_pixel][2] partial per vector column
  const int vpp = a.C / 8;        // 16-byte vectors per pixel
  const int cpg8 = (a.C / a.G) / 8;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * GN_PIX_PER_CTA, p1 = min(a.HW, p0 + GN_PIX_PER_CTA);
  for (int i = threadIdx.x; i < vpp * 2; i += GN_THREADS) sm[i] = 0.f;
  __syncthreads();
  // thread -> (vector column v, pixel lane): consecutive threads walk consecutive vectors of a pixel (coalesced)
  const int total = (p1 - p0) * vpp;
  float s = 0.f, ss = 0.f;
  int my_v = -1;
  int i_begin = threadIdx.x;
  if (GN_THREADS % vpp == 0 && !x2) {
    // every thread keeps ONE vector column and walks the pixels with a fixed stride: four independent 16-byte loads in
    // flight per thread (the one-load-per-iteration loop below left the memory pipe idle: 1.9 TB/s at the level-0 shape)
    const int v = threadIdx.x % vpp, step = GN_THREADS / vpp;
    const __nv_bfloat16* base = x + static_cast<long long>(b) * a.HW * a.ldx + v * 8;
    int p = p0 + threadIdx.x / vpp;
    for (; p + 3 * step < p1; p += 4 * step) {
      float f0[8], f1[8], f2[8], f3[8];
      ld8(base + static_cast<long long>(p) * a.ldx, f0);
      ld8(base + static_cast<long long>(p + step) * a.ldx, f1);
      ld8(base + static_cast<long long>(p + 2 * step) * a.ldx, f2);
      ld8(base + static_cast<long long>(p + 3 * step) * a.ldx, f3);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s += (f0[e] + f1[e]) + (f2[e] + f3[e]);
        ss += (f0[e] * f0[e] + f1[e] * f1[e]) + (f2[e] * f2[e] + f3[e] * f3[e]);
      }
    }
    my_v = v;
    i_begin = (p - p0) * vpp + v;       // the remaining pixels of this column go through the generic loop
  }
  for (int i = i_begin; i < total; i += GN_THREADS) {
    const int v = i % vpp, p = p0 + i / vpp;
    if (my_v >= 0 && v != my_v) {   // only when GN_THREADS % vpp != 0: flush and switch column
      atomicAdd(&sm[2 * my_v], s); atomicAdd(&sm[2 * my_v + 1], ss); s = ss = 0.f;
    }
    my_v = v;
    float f[8];
    ld8(x + (static_cast<long long>(b) * a.HW + p) * a.ldx + v * 8, f);
    if (x2) {
      float f2[8];
      ld8(x2 + (static_cast<long long>(b) * a.HW + p) * a.ldx2 + v * 8, f2);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16(f[e] + f2[e]));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s += f[e]; ss += f[e] * f[e]; }
  }
  if (my_v >= 0) { atomicAdd(&sm[2 * my_v], s); atomicAdd(&sm[2 * my_v + 1], ss); }
  __syncthreads();
  for (int g = threadIdx.x; g < a.G; g += GN_THREADS) {
    float gs = 0.f, gss = 0.f;
    for (int j = 0; j < cpg8; ++j) { gs += sm[2 * (g * cpg8 + j)]; gss += sm[2 * (g * cpg8 + j) + 1]; }
    atomicAdd(&stats[(static_cast<long long>(b) * a.G + g) * 2], static_cast<double>(gs));
    atomicAdd(&stats[(static_cast<long long>(b) * a.G + g) * 2 + 1], static_cast<double>(gss));
  }
  // The LAST CTA to finish turns the fp64 sums into the fp32 (mean, rstd) table the streaming kernels read (a separate
  // finalize launch sat on the critical chain of every GroupNorm: 96 two-microsecond kernels per step).
  __shared__ bool last_cta;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_cta = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last_cta) {
    __threadfence();
    for (int i = threadIdx.x; i < a.B * a.G; i += GN_THREADS) {
      const double m = __ldcg(&stats[2 * i]) / n_elem;
      double var = __ldcg(&stats[2 * i + 1]) / n_elem - m * m;
      var = var < 0.0 ? 0.0 : var;
      mr[i] = make_float2(static_cast<float>(m), static_cast<float>(1.0 / sqrt(var + static_cast<double>(a.eps))));
    }
  }
}

// (sum, sumsq) in fp64 -> (mean, rstd) in fp32 for statistics that a producer's epilogue accumulated
__global__ void gn_finalize_kernel(const double* __restrict__ stats, int BG, double n, float eps, float2* __restrict__ mr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BG) return;
  const double m = stats[2 * i] / n;
  double var = stats[2 * i + 1] / n - m * m;
  var = var < 0.0 ? 0.0 : var;
  mr[i] = make_float2(static_cast<float>(m), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
}
__device__ __forceinline__ void mean_rstd(const float2* mr, int b, int g, const GnArgs& a, float* mean, float* rstd) {
  const float2 v = __ldg(mr + static_cast<long long>(b) * a.G + g);
  *mean = v.x;
  *rstd = v.y;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// y = relu?((x (+x2) - mean) * rstd * gamma + beta)
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x2, const GnArgs a,
                const float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                __nv_bfloat16* __restrict__ y) {
  const int vpp = a.C / 8, cpg8 = (a.C / a.G) / 8;
  const long long total = static_cast<long long>(a.B) * a.HW * vpp;
  // the launcher makes gridDim.x * GN_THREADS a multiple of vpp, so every thread keeps ONE vector column: its
  // gamma / beta live in registers
  float gm[8], bt[8];
  {
    const int v0 = static_cast<int>((static_cast<long long>(blockIdx.x) * GN_THREADS + threadIdx.x) % vpp);
    load8f(gamma + v0 * 8, gm);
    load8f(beta + v0 * 8, bt);
  }
  for (long long i = static_cast<long long>(blockIdx.x) * GN_THREADS + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * GN_THREADS) {
    const int v = static_cast<int>(i % vpp);
    const long long px = i / vpp;
    const int b = static_cast<int>(px / a.HW);
    float mean, rstd;
    mean_rstd(stats, b, v / cpg8, a, &mean, &rstd);
    float f[8];
    ld8(x + px * a.ldx + v * 8, f);
    if (x2) {
      float f2[8];
      ld8(x2 + px * a.ldx2 + v * 8, f2);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16(f[e] + f2[e]));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float o = (f[e] - mean) * rstd * gm[e] + bt[e];
      f[e] = a.relu ? fmaxf(o, 0.f) : o;
    }
    st8(y + px * a.ldy + v * 8, f);
  }
}

// Backward pass 1: per (b,g): s1 = sum dy'*gamma, s2 = sum dy'*gamma*xhat;  per channel: dgamma += dy'*xhat,
// dbeta += dy'   (dy' = dy * [y > 0] when relu).  bstats (double, pre-zeroed) [B][G][2]; dgamma/dbeta fp32 pre-zeroed.
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x2,
                    const __nv_bfloat16* __restrict__ dy, long long lddy, const GnArgs a,
                    const float2* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                    double* __restrict__ bstats, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    unsigned* __restrict__ ticket, float2* __restrict__ s12) {
  extern __shared__ float sm[];   // per channel: dgamma, dbeta [2*C]; per vector column: s1, s2 [2*vpp]
  const int vpp = a.C / 8, cpg8 = (a.C / a.G) / 8;
  float* sm_dg = sm;
  float* sm_db = sm + a.C;
  float* sm_s = sm + 2 * a.C;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * GN_PIX_PER_CTA, p1 = min(a.HW, p0 + GN_PIX_PER_CTA);
  for (int i = threadIdx.x; i < 2 * a.C + 2 * vpp; i += GN_THREADS) sm[i] = 0.f;
  __syncthreads();
  const int total = (p1 - p0) * vpp;
  float dg[8], db[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) dg[e] = db[e] = 0.f;
  int my_v = -1;
  float gmr[8], btr[8];
  int loaded_v = -1;
  auto flush = [&]() {
#pragma unroll
    for (int e = 0; e < 8; ++e) { atomicAdd(&sm_dg[my_v * 8 + e], dg[e]); atomicAdd(&sm_db[my_v * 8 + e], db[e]); dg[e] = db[e] = 0.f; }
    atomicAdd(&sm_s[2 * my_v], s1); atomicAdd(&sm_s[2 * my_v + 1], s2); s1 = s2 = 0.f;
  };
  int i_begin = threadIdx.x;
  if (GN_THREADS % vpp == 0) {
    // fixed vector column per thread: group statistics / affine parameters are loaded once, and the x / dy vectors of
    // two pixels are requested together (4-6 independent 16-byte loads in flight per thread)
    const int v = threadIdx.x % vpp, step = GN_THREADS / vpp;
    float mean, rstd;
    mean_rstd(stats, b, v / cpg8, a, &mean, &rstd);
    load8f(gamma + v * 8, gmr);
    load8f(beta + v * 8, btr);
    loaded_v = v;
    int p = p0 + threadIdx.x / vpp;
    for (; p + step < p1; p += 2 * step) {
      const long long pa = static_cast<long long>(b) * a.HW + p, pb = pa + step;
      float fa[8], fb[8], da[8], db2[8];
      ld8(x + pa * a.ldx + v * 8, fa);
      ld8(x + pb * a.ldx + v * 8, fb);
      ld8(dy + pa * lddy + v * 8, da);
      ld8(dy + pb * lddy + v * 8, db2);
      if (x2) {
        float ga[8], gb[8];
        ld8(x2 + pa * a.ldx2 + v * 8, ga);
        ld8(x2 + pb * a.ldx2 + v * 8, gb);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          fa[e] = __bfloat162float(__float2bfloat16(fa[e] + ga[e]));
          fb[e] = __bfloat162float(__float2bfloat16(fb[e] + gb[e]));
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gm = gmr[e];
        const float xa = (fa[e] - mean) * rstd, xb = (fb[e] - mean) * rstd;
        float ya = da[e], yb = db2[e];
        if (a.relu && !(xa * gm + btr[e] > 0.f)) ya = 0.f;
        if (a.relu && !(xb * gm + btr[e] > 0.f)) yb = 0.f;
        dg[e] += ya * xa + yb * xb;
        db[e] += ya + yb;
        s1 += (ya + yb) * gm;
        s2 += (ya * xa + yb * xb) * gm;
      }
    }
    my_v = v;
    i_begin = (p - p0) * vpp + v;
  }
  for (int i = i_begin; i < total; i += GN_THREADS) {
    const int v = i % vpp, p = p0 + i / vpp;
    if (my_v >= 0 && v != my_v) flush();
    my_v = v;
    const long long px = static_cast<long long>(b) * a.HW + p;
    float mean, rstd;
    mean_rstd(stats, b, v / cpg8, a, &mean, &rstd);
    float f[8], d[8];
    ld8(x + px * a.ldx + v * 8, f);
    if (x2) {
      float f2[8];
      ld8(x2 + px * a.ldx2 + v * 8, f2);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16(f[e] + f2[e]));
    }
    ld8(dy + px * lddy + v * 8, d);
    if (loaded_v != v) { load8f(gamma + v * 8, gmr); load8f(beta + v * 8, btr); loaded_v = v; }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gm = gmr[e];
      const float xh = (f[e] - mean) * rstd;
      float dd = d[e];
      if (a.relu && !(xh * gm + btr[e] > 0.f)) dd = 0.f;
      dg[e] += dd * xh;
      db[e] += dd;
      s1 += dd * gm;
      s2 += dd * gm * xh;
    }
  }
  if (my_v >= 0) flush();
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += GN_THREADS) {
    atomicAdd(&dgamma[c], sm_dg[c]);
    atomicAdd(&dbeta[c], sm_db[c]);
  }
  for (int g = threadIdx.x; g < a.G; g += GN_THREADS) {
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < cpg8; ++j) { t1 += sm_s[2 * (g * cpg8 + j)]; t2 += sm_s[2 * (g * cpg8 + j) + 1]; }
    atomicAdd(&bstats[(static_cast<long long>(b) * a.G + g) * 2], static_cast<double>(t1));
    atomicAdd(&bstats[(static_cast<long long>(b) * a.G + g) * 2 + 1], static_cast<double>(t2));
  }
  // last CTA: fp64 sums -> fp32 (s1, s2) table (see gn_stats_kernel)
  __shared__ bool last_cta;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last_cta = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (last_cta) {
    __threadfence();
    for (int i = threadIdx.x; i < a.B * a.G; i += GN_THREADS)
      s12[i] = make_float2(static_cast<float>(__ldcg(&bstats[2 * i])), static_cast<float>(__ldcg(&bstats[2 * i + 1])));
  }
}

// Backward pass 2: dx = rstd * (dy'*gamma - (s1 + xhat*s2)/n)
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ x2,
                    const __nv_bfloat16* __restrict__ dy, long long lddy, const GnArgs a,
                    const float2* __restrict__ stats, const float2* __restrict__ bstats,
                    const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ dx,
                    long long lddx, float* __restrict__ dx_colsum) {
  // dx_colsum (optional, fp32 [C], ADDED to): per-channel sum of dx over all pixels = the bias gradient of the layer that
  // produced x (a conv / DCN bias directly in front of this norm) -- this kernel holds every dx value in registers
  // anyway, the separate read-only column-sum pass over dx disappears
  extern __shared__ float sm_cs[];
  const int vpp = a.C / 8, cpg8 = (a.C / a.G) / 8;
  const float inv_n = 1.f / (static_cast<float>(a.HW) * (a.C / a.G));
  const long long total = static_cast<long long>(a.B) * a.HW * vpp;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (dx_colsum) {
    for (int i = threadIdx.x; i < a.C; i += GN_THREADS) sm_cs[i] = 0.f;
    __syncthreads();
  }
  float gmr[8], btr[8];
  {
    const int v0 = static_cast<int>((static_cast<long long>(blockIdx.x) * GN_THREADS + threadIdx.x) % vpp);
    load8f(gamma + v0 * 8, gmr);
    load8f(beta + v0 * 8, btr);
  }
  for (long long i = static_cast<long long>(blockIdx.x) * GN_THREADS + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * GN_THREADS) {
    const int v = static_cast<int>(i % vpp);
    const long long px = i / vpp;
    const int b = static_cast<int>(px / a.HW), g = v / cpg8;
    float mean, rstd;
    mean_rstd(stats, b, g, a, &mean, &rstd);
    const float2 s12 = __ldg(bstats + static_cast<long long>(b) * a.G + g);
    const float s1 = s12.x, s2 = s12.y;
    float f[8], d[8];
    ld8(x + px * a.ldx + v * 8, f);
    if (x2) {
      float f2[8];
      ld8(x2 + px * a.ldx2 + v * 8, f2);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16(f[e] + f2[e]));
    }
    ld8(dy + px * lddy + v * 8, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gm = gmr[e];
      const float xh = (f[e] - mean) * rstd;
      float dd = d[e];
      if (a.relu && !(xh * gm + btr[e] > 0.f)) dd = 0.f;
      f[e] = rstd * (dd * gm - (s1 + xh * s2) * inv_n);
      cs[e] += f[e];
    }
    st8(dx + px * lddx + v * 8, f);
  }
  if (dx_colsum) {      // the launcher keeps ONE vector column per thread: fold the block's threads of a column, then one atomic
    const int v0 = static_cast<int>((static_cast<long long>(blockIdx.x) * GN_THREADS + threadIdx.x) % vpp);
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&sm_cs[v0 * 8 + e], cs[e]);
    __syncthreads();
    for (int i = threadIdx.x; i < a.C; i += GN_THREADS) atomicAdd(&dx_colsum[i], sm_cs[i]);
  }
}

static int gn_check(const char* who, int C, int G, long long ldx, long long ldy) {
  if (G < 1 || C % G || (C / G) % 8 || (ldx % 8) || (ldy % 8) || C > 4096)
    return set_error("%s: needs (C/G) %% 8 == 0 and 16-byte aligned pitches (C=%d G=%d)", who, C, G);
  return 0;
}
// grid-stride launches: gridDim.x * GN_THREADS must be a multiple of vpp (= C/8) so that a thread keeps its column
static int gn_grid(long long total_vec, int vpp) {
  long long blocks = (total_vec + GN_THREADS - 1) / GN_THREADS;
  long long g = blocks < 148 * 8 ? (blocks < 1 ? 1 : blocks) : 148 * 8;
  while ((g * GN_THREADS) % vpp) ++g;
  return static_cast<int>(g);
}

}  // namespace lsn

using namespace lsn;

extern "C" int lsnet_groupnorm_fwd(const void* x, long long ldx, const void* x2, long long ldx2, int B, int HW, int C,
                                   int G, const float* gamma, const float* beta, float eps, int relu, double* stats,
                                   void* y, long long ldy, void* stream) {
  if (B <= 0 || HW <= 0) return 0;
  if (int rc = gn_check("lsnet_groupnorm_fwd", C, G, ldx, ldy)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GnArgs a{B, HW, C, G, ldx, ldx2, ldy, eps, relu};
  // workspace layout (3*B*G + 1 doubles): [2*B*G fp64 sums][ticket counter][B*G float2 (mean, rstd)]
  cudaMemsetAsync(stats, 0, sizeof(double) * (2 * B * G + 1), st);
  dim3 grid((HW + GN_PIX_PER_CTA - 1) / GN_PIX_PER_CTA, B);
  float2* mr = reinterpret_cast<float2*>(stats + 2 * B * G + 1);
  gn_stats_kernel<<<grid, GN_THREADS, sizeof(float) * 2 * (C / 8), st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(x2), a, stats,
      reinterpret_cast<unsigned*>(stats + 2 * B * G), mr, static_cast<double>(HW) * (C / G));
  if (int rc = check_launch("gn_stats")) return rc;
  gn_apply_kernel<<<gn_grid(static_cast<long long>(B) * HW * (C / 8), C / 8), GN_THREADS, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(x2), a, mr, gamma, beta,
      static_cast<__nv_bfloat16*>(y));
  return check_launch("gn_apply");
}

extern "C" int lsnet_groupnorm_fwd_pre(const void* x, long long ldx, int B, int HW, int C, int G, const float* gamma,
                                       const float* beta, float eps, int relu, double* stats, void* y, long long ldy,
                                       void* stream) {
  if (B <= 0 || HW <= 0) return 0;
  if (int rc = gn_check("lsnet_groupnorm_fwd_pre", C, G, ldx, ldy)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GnArgs a{B, HW, C, G, ldx, 0, ldy, eps, relu};
  float2* mr = reinterpret_cast<float2*>(stats + 2 * B * G + 1);
  gn_finalize_kernel<<<(B * G + 127) / 128, 128, 0, st>>>(stats, B * G, static_cast<double>(HW) * (C / G), eps, mr);
  if (int rc = check_launch("gn_finalize")) return rc;
  gn_apply_kernel<<<gn_grid(static_cast<long long>(B) * HW * (C / 8), C / 8), GN_THREADS, 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), nullptr, a, mr, gamma, beta, static_cast<__nv_bfloat16*>(y));
  return check_launch("gn_apply");
}

static int groupnorm_bwd_impl(const void* x, long long ldx, const void* x2, long long ldx2, const void* dy,
                              long long lddy, int B, int HW, int C, int G, const float* gamma, const float* beta,
                              float eps, int relu, const double* stats, double* ws_bstats, void* dx, long long lddx,
                              float* dgamma, float* dbeta, int accumulate, float* dx_colsum, void* stream) {
  if (B <= 0 || HW <= 0) return 0;
  if (int rc = gn_check("lsnet_groupnorm_bwd", C, G, ldx, lddx)) return rc;
  if (lddy % 8) return set_error("lsnet_groupnorm_bwd: dy pitch must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GnArgs a{B, HW, C, G, ldx, ldx2, 0, eps, relu};
  cudaMemsetAsync(ws_bstats, 0, sizeof(double) * (2 * B * G + 1), st);
  if (!accumulate) {
    cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st);
  }
  dim3 grid((HW + GN_PIX_PER_CTA - 1) / GN_PIX_PER_CTA, B);
  const float2* mr = reinterpret_cast<const float2*>(stats + 2 * B * G + 1);
  float2* s12 = reinterpret_cast<float2*>(ws_bstats + 2 * B * G + 1);
  gn_bwd_stats_kernel<<<grid, GN_THREADS, sizeof(float) * (2 * C + 2 * (C / 8)), st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(x2),
      static_cast<const __nv_bfloat16*>(dy), lddy, a, mr, gamma, beta, ws_bstats, dgamma, dbeta,
      reinterpret_cast<unsigned*>(ws_bstats + 2 * B * G), s12);
  if (int rc = check_launch("gn_bwd_stats")) return rc;
  // with the column sums every block ends with C global reds onto the same C addresses: half the blocks (still four per
  // SM) keep that tail short
  int apply_grid = gn_grid(static_cast<long long>(B) * HW * (C / 8), C / 8);
  if (dx_colsum && apply_grid > 148 * 4) {
    apply_grid = 148 * 4;
    while ((static_cast<long long>(apply_grid) * GN_THREADS) % (C / 8)) ++apply_grid;
  }
  gn_bwd_apply_kernel<<<apply_grid, GN_THREADS, dx_colsum ? sizeof(float) * C : 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(x2),
      static_cast<const __nv_bfloat16*>(dy), lddy, a, mr, s12, gamma, beta, static_cast<__nv_bfloat16*>(dx),
      lddx, dx_colsum);
  return check_launch("gn_bwd_apply");
}

extern "C" int lsnet_groupnorm_bwd(const void* x, long long ldx, const void* x2, long long ldx2, const void* dy,
                                   long long lddy, int B, int HW, int C, int G, const float* gamma, const float* beta,
                                   float eps, int relu, const double* stats, double* ws_bstats, void* dx, long long lddx,
                                   float* dgamma, float* dbeta, void* stream) {
  return groupnorm_bwd_impl(x, ldx, x2, ldx2, dy, lddy, B, HW, C, G, gamma, beta, eps, relu, stats, ws_bstats, dx, lddx,
                            dgamma, dbeta, 0, nullptr, stream);
}
// same, but dgamma / dbeta are ADDED to (the affine parameters' gradient memory; no memset, no separate add kernel)
extern "C" int lsnet_groupnorm_bwd_acc(const void* x, long long ldx, const void* x2, long long ldx2, const void* dy,
                                       long long lddy, int B, int HW, int C, int G, const float* gamma, const float* beta,
                                       float eps, int relu, const double* stats, double* ws_bstats, void* dx,
                                       long long lddx, float* dgamma, float* dbeta, float* dx_colsum, void* stream) {
  return groupnorm_bwd_impl(x, ldx, x2, ldx2, dy, lddy, B, HW, C, G, gamma, beta, eps, relu, stats, ws_bstats, dx, lddx,
                            dgamma, dbeta, 1, dx_colsum, stream);
}
