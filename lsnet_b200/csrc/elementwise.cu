// Small streaming helpers of the backward pass (sm_100a), fused so that every upstream gradient is read exactly once:
//   grad_prep   dY (fp32 or bf16, any pixel pitch, C channels) -> bf16 pixel-major buffer with the channel count padded
//               to the GEMM's alignment, optionally masked by the ReLU of the forward output (conv+ReLU epilogues),
//               and -- in the same pass -- the per-channel column sum = bias gradient (replaces the reference's
//               `grad_bias.addmv_(grad_output, ones)` GEMV, deform_conv_cuda.cpp:788-794, and torch's separate
//               cast / mask / sum kernels).
#include <stdint.h>

#include "common.cuh"
#include "lsnet_internal.h"

namespace lsn {

constexpr int PREP_THREADS = 256;
constexpr int PREP_ROWS = 64;   // rows per CTA

template <bool FP32_IN>
__global__ void __launch_bounds__(PREP_THREADS)
grad_prep_kernel(const void* __restrict__ gy, long long ldg, const __nv_bfloat16* __restrict__ relu_out, long long ldo,
                 long long P, int C, int Cpad, __nv_bfloat16* __restrict__ out, long long ldout,
                 float* __restrict__ colsum) {
  extern __shared__ float sm[];   // [Cpad] partial column sums
  for (int i = threadIdx.x; i < Cpad; i += PREP_THREADS) sm[i] = 0.f;
  __syncthreads();
  const long long r0 = static_cast<long long>(blockIdx.x) * PREP_ROWS;
  const long long r1 = min(P, r0 + PREP_ROWS);
  // thread -> fixed column (stride PREP_THREADS over the row-major tile keeps columns fixed when Cpad | PREP_THREADS)
  const long long total = (r1 - r0) * Cpad;
  float acc = 0.f;
  int my_c = -1;
  for (long long i = threadIdx.x; i < total; i += PREP_THREADS) {
    const int c = static_cast<int>(i % Cpad);
    const long long r = r0 + i / Cpad;
    if (colsum && my_c >= 0 && c != my_c) { atomicAdd(&sm[my_c], acc); acc = 0.f; }
    my_c = c;
    float v = 0.f;
    if (c < C) {
      v = FP32_IN ? static_cast<const float*>(gy)[r * ldg + c]
                  : __bfloat162float(static_cast<const __nv_bfloat16*>(gy)[r * ldg + c]);
      if (relu_out && !(__bfloat162float(relu_out[r * ldo + c]) > 0.f)) v = 0.f;
    }
    out[r * ldout + c] = __float2bfloat16(v);
    acc += v;
  }
  if (colsum) {
    if (my_c >= 0) atomicAdd(&sm[my_c], acc);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += PREP_THREADS)
      if (sm[i] != 0.f) atomicAdd(&colsum[i], sm[i]);
  }
}

// Vector path (C % 8 == 0, Cpad == C, 16-byte aligned pitches, (C/8) a power of two): one thread = one 16-byte vector of
// 8 channels, walking rows with a fixed column, so the column sums live in registers until the end.  A CTA covers a
// slice of at most 256 channels (grid.y) x PREPV_RPT rows per row-lane (grid.x), so wide maps (the 1024 / 2048-channel
// trunk stages) still fill the machine, and every thread keeps PREPV_U independent 16-byte loads in flight.
constexpr int PREPV_RPT = 8;     // rows per thread
constexpr int PREPV_U = 4;       // loads in flight per thread
template <bool FP32_IN>
__global__ void __launch_bounds__(PREP_THREADS)
grad_prep_vec_kernel(const void* __restrict__ gy, long long ldg, const __nv_bfloat16* __restrict__ relu_out,
                     long long ldo, long long P, int C, __nv_bfloat16* __restrict__ out, long long ldout,
                     float* __restrict__ colsum) {
  __shared__ float sm[256];        // column sums of this CTA's channel slice
  const int vpp = C / 8;
  const int vpc = vpp < 32 ? vpp : 32;              // vectors per row handled by this CTA
  const int lanes = PREP_THREADS / vpc;             // row lanes
  const int v = blockIdx.y * vpc + threadIdx.x % vpc;
  const int rl = threadIdx.x / vpc;
  if (colsum) {
    for (int i = threadIdx.x; i < 256; i += PREP_THREADS) sm[i] = 0.f;
    __syncthreads();
  }
  const long long r0 = static_cast<long long>(blockIdx.x) * lanes * PREPV_RPT + rl;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < PREPV_RPT; k0 += PREPV_U) {
    float f[PREPV_U][8];
    uint4 mk[PREPV_U];
#pragma unroll
    for (int u = 0; u < PREPV_U; ++u) {
      const long long r = r0 + static_cast<long long>(k0 + u) * lanes;
      const bool ok = r < P;
      if (FP32_IN) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (ok) {
          a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(gy) + r * ldg + v * 8));
          b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(gy) + r * ldg + v * 8 + 4));
        }
        f[u][0] = a.x; f[u][1] = a.y; f[u][2] = a.z; f[u][3] = a.w; f[u][4] = b.x; f[u][5] = b.y; f[u][6] = b.z; f[u][7] = b.w;
      } else {
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (ok) q = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(gy) + r * ldg + v * 8));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[u][2 * i] = t.x; f[u][2 * i + 1] = t.y; }
      }
      mk[u] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);     // bf16 1.0: keep everything
      if (relu_out && ok) mk[u] = __ldg(reinterpret_cast<const uint4*>(relu_out + r * ldo + v * 8));
    }
#pragma unroll
    for (int u = 0; u < PREPV_U; ++u) {
      const long long r = r0 + static_cast<long long>(k0 + u) * lanes;
      if (r >= P) continue;
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&mk[u]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        if (!(t.x > 0.f)) f[u][2 * i] = 0.f;
        if (!(t.y > 0.f)) f[u][2 * i + 1] = 0.f;
      }
      if (out)     // out == nullptr: column sums only (the gradient is already in the layout the GEMMs consume)
        *reinterpret_cast<uint4*>(out + r * ldout + v * 8) =
            make_uint4(pack_bf16x2(f[u][0], f[u][1]), pack_bf16x2(f[u][2], f[u][3]), pack_bf16x2(f[u][4], f[u][5]),
                       pack_bf16x2(f[u][6], f[u][7]));
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[u][e];
    }
  }
  if (colsum) {
    const int cl = (threadIdx.x % vpc) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&sm[cl + e], acc[e]);
    __syncthreads();
    for (int i = threadIdx.x; i < vpc * 8; i += PREP_THREADS)
      if (sm[i] != 0.f) atomicAdd(&colsum[blockIdx.y * vpc * 8 + i], sm[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Frozen-statistics BatchNorm folded into the preceding convolution's weights (ResNet trunk, norm_eval=True:
// mmdet/models/backbones/resnet.py:636-646):  W'[o] = W[o] * s[o],  b'[o] = beta[o] - mean[o] * s[o],
// s = gamma / sqrt(var + eps).  One block per output channel; W is OIHW fp32, W' is written bf16 in OHWI
// (channels_last) order for cuDNN's NHWC kernels.  Backward: gW = gW' * s, ggamma = (sum gW'.W - gb'.mean)/sqrt(var+eps),
// gbeta = gb'.
// ---------------------------------------------------------------------------------------------------------------
__global__ void bn_fold_fwd_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ mean,
                                   const float* __restrict__ var, float eps, int I, int KK,
                                   __nv_bfloat16* __restrict__ Wb, float* __restrict__ bias) {
  const int o = blockIdx.x;
  const float s = gamma[o] * rsqrtf(var[o] + eps);
  const int n = I * KK;
  const float* w = W + static_cast<long long>(o) * n;
  __nv_bfloat16* d = Wb + static_cast<long long>(o) * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {   // j indexes the OHWI destination
    const int k = j / I, i = j % I;
    d[j] = __float2bfloat16(w[i * KK + k] * s);
  }
  if (threadIdx.x == 0) bias[o] = beta[o] - mean[o] * s;
}

__global__ void bn_fold_bwd_kernel(const __nv_bfloat16* __restrict__ gWb, const float* __restrict__ gbias,
                                   const float* __restrict__ W, const float* __restrict__ gamma,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps, int I,
                                   int KK, float* __restrict__ gW, float* __restrict__ ggamma,
                                   float* __restrict__ gbeta) {
  __shared__ float red[32];
  const int o = blockIdx.x;
  const float r = rsqrtf(var[o] + eps);
  const float s = gamma[o] * r;
  const int n = I * KK;
  const float* w = W + static_cast<long long>(o) * n;
  const __nv_bfloat16* g = gWb + static_cast<long long>(o) * n;
  float* d = gW + static_cast<long long>(o) * n;
  float acc = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int k = j / I, i = j % I;
    const float gv = __bfloat162float(g[j]);
    const float wv = w[i * KK + k];
    d[i * KK + k] = gv * s;
    acc += gv * wv;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      const float gb = gbias ? gbias[o] : 0.f;
      ggamma[o] = (v - gb * mean[o]) * r;
      gbeta[o] = gb;
    }
  }
}


// ---- GEMM-operand variant of the fold (own tcgen05 trunk convolutions) -------------------------------------------
// W fp32 with element (o, i, k) at o*I*KK + i*si + k*sk (OIHW: si = KK, sk = 1; tap-major OHWI: si = 1, sk = I).
// forward : Wb[o][k][i] = W*s[o]  (bf16, A-side pack [O, KK*I]),  Wt[i][k][o] = the same value (bf16, pack of the input-
//           gradient GEMM [I, KK*O]),  bias[o] = beta[o] - mean[o]*s[o].  32 x 32 (o, i) tiles through shared memory so
//           that both packs are written with unit stride.
// backward: gW(o,i,k) (+)= gWb[o][k][i]*s[o]  (gWb fp32 [O, KK*I] straight from the weight-gradient GEMM),
//           ggamma[o] (+)= (sum gWb*W - gbias*mean)*r,  gbeta[o] (+)= gbias[o].
__global__ void bn_fold2_fwd_kernel(const float* __restrict__ W, long long si, long long sk,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps, int O, int I,
                                    int KK, __nv_bfloat16* __restrict__ Wb, __nv_bfloat16* __restrict__ Wt,
                                    float* __restrict__ bias) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32, k = blockIdx.z;
  const int tx = threadIdx.x, ty = threadIdx.y;      // (32, 8)
  const bool i_fast = si == 1;
  for (int r = ty; r < 32; r += 8) {
    // read with the unit-stride index on tx
    const int o = i_fast ? o0 + r : o0 + tx, i = i_fast ? i0 + tx : i0 + r;
    float v = 0.f;
    if (o < O && i < I) v = W[static_cast<long long>(o) * I * KK + i * si + k * sk] * (gamma[o] * rsqrtf(var[o] + eps));
    if (i_fast) tile[r][tx] = v; else tile[tx][r] = v;      // tile[o - o0][i - i0]
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int o = o0 + r, i = i0 + tx;
    if (o < O && i < I) Wb[(static_cast<long long>(o) * KK + k) * I + i] = __float2bfloat16(tile[r][tx]);
    const int i2 = i0 + r, o2 = o0 + tx;
    if (o2 < O && i2 < I) Wt[(static_cast<long long>(i2) * KK + k) * O + o2] = __float2bfloat16(tile[tx][r]);
  }
  if (blockIdx.x == 0 && k == 0 && ty == 0 && o0 + tx < O) {
    const int o = o0 + tx;
    bias[o] = beta[o] - mean[o] * gamma[o] * rsqrtf(var[o] + eps);
  }
}

__global__ void bn_fold2_bwd_kernel(const float* __restrict__ gWb, const float* __restrict__ gbias,
                                    const float* __restrict__ W, long long si, long long sk,
                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, int I, int KK, float* __restrict__ gW,
                                    float* __restrict__ ggamma, float* __restrict__ gbeta, int accumulate) {
  __shared__ float red[32];
  const int o = blockIdx.x;
  const float r = rsqrtf(var[o] + eps);
  const float s = gamma[o] * r;
  const int n = I * KK;
  const float* w = W + static_cast<long long>(o) * n;
  const float* g = gWb + static_cast<long long>(o) * n;
  float* d = gW + static_cast<long long>(o) * n;
  float acc = 0.f;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {      // j = k*I + i
    const int k = j / I, i = j - k * I;
    const long long e = i * si + k * sk;
    const float gv = g[j];
    acc += gv * w[e];
    d[e] = accumulate ? d[e] + gv * s : gv * s;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      const float gb = gbias ? gbias[o] : 0.f;
      const float gg = (v - gb * mean[o]) * r;
      ggamma[o] = accumulate ? ggamma[o] + gg : gg;
      gbeta[o] = accumulate ? gbeta[o] + gb : gb;
    }
  }
}

}  // namespace lsn

using namespace lsn;

extern "C" int lsnet_bn_fold_fwd(const float* W, const float* gamma, const float* beta, const float* mean,
                                 const float* var, float eps, int O, int I, int KK, void* Wb, float* bias,
                                 void* stream) {
  if (O <= 0) return 0;
  bn_fold_fwd_kernel<<<O, 128, 0, static_cast<cudaStream_t>(stream)>>>(W, gamma, beta, mean, var, eps, I, KK,
                                                                      static_cast<__nv_bfloat16*>(Wb), bias);
  return check_launch("bn_fold_fwd");
}

extern "C" int lsnet_bn_fold_bwd(const void* gWb, const float* gbias, const float* W, const float* gamma,
                                 const float* mean, const float* var, float eps, int O, int I, int KK, float* gW,
                                 float* ggamma, float* gbeta, void* stream) {
  if (O <= 0) return 0;
  bn_fold_bwd_kernel<<<O, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(gWb), gbias, W,
                                                                      gamma, mean, var, eps, I, KK, gW, ggamma, gbeta);
  return check_launch("bn_fold_bwd");
}

extern "C" int lsnet_bn_fold2_fwd(const float* W, long long si, long long sk, const float* gamma, const float* beta,
                                  const float* mean, const float* var, float eps, int O, int I, int KK, void* Wb,
                                  void* Wt, float* bias, void* stream) {
  if (O <= 0) return 0;
  if (!((si == 1 && sk == I) || (si == KK && sk == 1)))
    return set_error("lsnet_bn_fold2_fwd: weight must be OIHW-contiguous or tap-major (si=%lld sk=%lld)", si, sk);
  dim3 grid((I + 31) / 32, (O + 31) / 32, KK);
  bn_fold2_fwd_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      W, si, sk, gamma, beta, mean, var, eps, O, I, KK, static_cast<__nv_bfloat16*>(Wb), static_cast<__nv_bfloat16*>(Wt), bias);
  return check_launch("bn_fold2_fwd");
}

extern "C" int lsnet_bn_fold2_bwd(const float* gWb, const float* gbias, const float* W, long long si, long long sk,
                                  const float* gamma, const float* mean, const float* var, float eps, int O, int I,
                                  int KK, float* gW, float* ggamma, float* gbeta, int accumulate, void* stream) {
  if (O <= 0) return 0;
  bn_fold2_bwd_kernel<<<O, 256, 0, static_cast<cudaStream_t>(stream)>>>(gWb, gbias, W, si, sk, gamma, mean, var, eps, I,
                                                                       KK, gW, ggamma, gbeta, accumulate);
  return check_launch("bn_fold2_bwd");
}

static int grad_prep_impl(const void* gy, int gy_fp32, long long ldg, const void* relu_out, long long ldo, long long P,
                          int C, int Cpad, void* out, long long ldout, float* colsum, int accumulate, void* stream) {
  if (P <= 0) return 0;
  if (Cpad < C || Cpad > 8192) return set_error("lsnet_grad_prep: bad channel counts C=%d Cpad=%d", C, Cpad);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (colsum && !accumulate) cudaMemsetAsync(colsum, 0, sizeof(float) * C, st);
  const int vpp = C / 8;
  const bool vec = (C % 8 == 0) && Cpad == C && vpp >= 1 && (vpp & (vpp - 1)) == 0 &&
                   (ldg % (gy_fp32 ? 4 : 8) == 0) && (!out || ldout % 8 == 0) && (!relu_out || ldo % 8 == 0) &&
                   (reinterpret_cast<uintptr_t>(gy) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                   (!relu_out || reinterpret_cast<uintptr_t>(relu_out) % 16 == 0);
  if (vec) {
    const int vpc = vpp < 32 ? vpp : 32;
    const int rows_per_cta = (PREP_THREADS / vpc) * PREPV_RPT;
    const dim3 gridv(static_cast<unsigned>((P + rows_per_cta - 1) / rows_per_cta), static_cast<unsigned>(vpp / vpc));
    if (gy_fp32)
      grad_prep_vec_kernel<true><<<gridv, PREP_THREADS, 0, st>>>(
          gy, ldg, static_cast<const __nv_bfloat16*>(relu_out), ldo, P, C, static_cast<__nv_bfloat16*>(out), ldout,
          colsum);
    else
      grad_prep_vec_kernel<false><<<gridv, PREP_THREADS, 0, st>>>(
          gy, ldg, static_cast<const __nv_bfloat16*>(relu_out), ldo, P, C, static_cast<__nv_bfloat16*>(out), ldout,
          colsum);
    return check_launch("grad_prep_vec");
  }
  if (!out) return set_error("lsnet_grad_prep: out == NULL (column sums only) needs the vector path");
  const int grid = static_cast<int>((P + PREP_ROWS - 1) / PREP_ROWS);
  if (gy_fp32)
    grad_prep_kernel<true><<<grid, PREP_THREADS, sizeof(float) * Cpad, st>>>(
        gy, ldg, static_cast<const __nv_bfloat16*>(relu_out), ldo, P, C, Cpad, static_cast<__nv_bfloat16*>(out), ldout,
        colsum);
  else
    grad_prep_kernel<false><<<grid, PREP_THREADS, sizeof(float) * Cpad, st>>>(
        gy, ldg, static_cast<const __nv_bfloat16*>(relu_out), ldo, P, C, Cpad, static_cast<__nv_bfloat16*>(out), ldout,
        colsum);
  return check_launch("grad_prep");
}

extern "C" int lsnet_grad_prep(const void* gy, int gy_fp32, long long ldg, const void* relu_out, long long ldo,
                               long long P, int C, int Cpad, void* out, long long ldout, float* colsum, void* stream) {
  return grad_prep_impl(gy, gy_fp32, ldg, relu_out, ldo, P, C, Cpad, out, ldout, colsum, 0, stream);
}
// same, but the column sums are ADDED to colsum (the bias parameter's gradient memory) instead of replacing it
extern "C" int lsnet_grad_prep_acc(const void* gy, int gy_fp32, long long ldg, const void* relu_out, long long ldo,
                                   long long P, int C, int Cpad, void* out, long long ldout, float* colsum, void* stream) {
  return grad_prep_impl(gy, gy_fp32, ldg, relu_out, ldo, P, C, Cpad, out, ldout, colsum, 1, stream);
}


// ---------------------------------------------------------------------------------------------------------------
// Optimizer step over the flat parameter / gradient / momentum buffers in ONE pass (mmcv OptimizerHook.after_train_iter,
// mmcv/runner/hooks/optimizer.py:19-28: clip_grad_norm_ then torch.optim.SGD with momentum, dampening 0, no nesterov):
//   s = min(1, max_norm / (|g| + 1e-6))          (|g| = *grad_norm, computed on the device; max_norm <= 0: no clip)
//   g' = s * g + wd * p ;  m = momentum * m + g' ;  p = p - lr * m
// HBM-bound: 12 bytes read + 8 bytes written per parameter.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, long long n4,
                    const float* __restrict__ grad_norm, float max_norm, float lr, float momentum, float wd) {
  float s = 1.f;
  if (max_norm > 0.f && grad_norm) s = fminf(1.f, max_norm / (__ldg(grad_norm) + 1e-6f));
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    mv.x = momentum * mv.x + (s * gv.x + wd * pv.x); pv.x -= lr * mv.x;
    mv.y = momentum * mv.y + (s * gv.y + wd * pv.y); pv.y -= lr * mv.y;
    mv.z = momentum * mv.z + (s * gv.z + wd * pv.z); pv.z -= lr * mv.z;
    mv.w = momentum * mv.w + (s * gv.w + wd * pv.w); pv.w -= lr * mv.w;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
}

extern "C" int lsnet_sgd_momentum_step(float* params, const float* grads, float* momentum_buf, long long n,
                                       const float* grad_norm, float max_norm, float lr, float momentum,
                                       float weight_decay, void* stream) {
  if (n <= 0) return 0;
  if ((n % 4) || (reinterpret_cast<uintptr_t>(params) % 16) || (reinterpret_cast<uintptr_t>(grads) % 16) ||
      (reinterpret_cast<uintptr_t>(momentum_buf) % 16))
    return set_error("lsnet_sgd_momentum_step: flat buffers must be 16-byte aligned with n %% 4 == 0 (n=%lld)", n);
  const long long n4 = n / 4;
  const int grid = static_cast<int>(n4 / 256 + 1 < 148LL * 8 ? n4 / 256 + 1 : 148LL * 8);
  sgd_momentum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, momentum_buf, n4, grad_norm,
                                                                           max_norm, lr, momentum, weight_decay);
  return check_launch("sgd_momentum");
}


// ---------------------------------------------------------------------------------------------------------------
// LSHead element-wise glue (lsnet_head.py:372-400, 502-598, 735-755) as two kernels each way instead of ~20 torch ops
// per level:
//   pred_reg:      sp = softplus(o[:, :n_sp])  (beta 1, threshold 20)  and the 18 DCN sampling offsets
//                  off[j] = signed_pair(sp[src_j], sp[src_j + 1]) - base[j]   (mode 0: '-' slot wins ties and is negated)
//                  off[j] = o[src_j] - base[j]                                 (mode 1: free offset channels of 'bbox')
//                  backward: d o = softplus'(o) * (d sp + gradient_mul * routed d off)   (the reference mixes
//                  (1 - gm) * reg.detach() + gm * reg, i.e. value 1x, gradient gm x)
//   add_softplus:  y = softplus(t + s) with s detached (refine = softplus(raw + init.detach()))
// ---------------------------------------------------------------------------------------------------------------
constexpr int PR_MAX_OUT = 160;
constexpr int PR_MAX_OFF = 32;
struct PredRegTab {
  int n_off;                        // DCN offset channels (2 * kernel points)
  short src[PR_MAX_OFF];            // first source channel of offset j
  signed char mode[PR_MAX_OFF];     // 0: signed pair (src, src+1) of sp; 1: raw channel src
  float base[PR_MAX_OFF];           // dcn_base_offset
  signed char inv_j[PR_MAX_OUT];    // for every channel of o: the offset it feeds (-1: none) ...
  signed char inv_slot[PR_MAX_OUT]; // ... and its slot in the pair (0 / 1; 2: raw)
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float softplus_grad_f(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
pred_reg_fwd_kernel(const float* __restrict__ o, long long ldo, long long P, int n_sp, const PredRegTab tab,
                    float* __restrict__ sp, long long ldsp, float* __restrict__ off, long long ldoff) {
  const long long n1 = P * n_sp, n2 = P * tab.n_off;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n1 + n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < n1) {
      const long long p = i / n_sp;
      const int c = static_cast<int>(i - p * n_sp);
      sp[p * ldsp + c] = softplus_f(__ldg(o + p * ldo + c));
    } else {
      const long long k = i - n1, p = k / tab.n_off;
      const int j = static_cast<int>(k - p * tab.n_off);
      const float* row = o + p * ldo + tab.src[j];
      float v;
      if (tab.mode[j] == 0) {
        const float a = softplus_f(__ldg(row)), b = softplus_f(__ldg(row + 1));
        v = (a >= b) ? -a : b;      // torch.max returns the first maximum: the '-' slot wins ties
      } else {
        v = __ldg(row);
      }
      off[p * ldoff + j] = v - tab.base[j];
    }
  }
}

__global__ void __launch_bounds__(256)
pred_reg_bwd_kernel(const float* __restrict__ o, long long ldo, long long P, int n_sp, int n_out, const PredRegTab tab,
                    const float* __restrict__ gsp, long long ldgsp, const float* __restrict__ goff, long long ldgoff,
                    float gmul, float* __restrict__ go, long long ldgo) {
  const long long n = P * n_out;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / n_out;
    const int c = static_cast<int>(i - p * n_out);
    float g = (c < n_sp && gsp) ? __ldg(gsp + p * ldgsp + c) : 0.f;
    const int j = tab.inv_j[c];
    if (j >= 0 && goff) {
      const float gj = gmul * __ldg(goff + p * ldgoff + j);
      const int slot = tab.inv_slot[c];
      if (slot == 2) {
        g += gj;
      } else {
        const float* row = o + p * ldo + tab.src[j];
        const float a = softplus_f(__ldg(row)), b = softplus_f(__ldg(row + 1));
        const int sel = (a >= b) ? 0 : 1;
        if (slot == sel) g += sel == 0 ? -gj : gj;
      }
    }
    if (c < n_sp) g *= softplus_grad_f(__ldg(o + p * ldo + c));
    go[p * ldgo + c] = g;
  }
}

template <bool BWD>
__global__ void __launch_bounds__(256)
add_softplus_kernel(const float* __restrict__ t, long long ldt, const float* __restrict__ s, long long lds,
                    const float* __restrict__ gy, long long ldgy, long long P, int C, float* __restrict__ out,
                    long long ldout) {
  const long long n = P * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i / C;
    const int c = static_cast<int>(i - p * C);
    const float x = __ldg(t + p * ldt + c) + __ldg(s + p * lds + c);
    out[p * ldout + c] = BWD ? __ldg(gy + p * ldgy + c) * softplus_grad_f(x) : softplus_f(x);
  }
}

static int ew_grid(long long n) {
  const long long b = (n + 255) / 256;
  return static_cast<int>(b < 148LL * 16 ? (b < 1 ? 1 : b) : 148LL * 16);
}

static int fill_tab(PredRegTab* tab, int n_off, const int* src, const int* mode, const float* base, int n_sp, int n_out) {
  if (n_off < 1 || n_off > PR_MAX_OFF || n_out > PR_MAX_OUT || n_sp > n_out || n_sp < 0)
    return set_error("lsnet_pred_reg: unsupported sizes (n_off=%d n_sp=%d n_out=%d)", n_off, n_sp, n_out);
  tab->n_off = n_off;
  for (int c = 0; c < PR_MAX_OUT; ++c) { tab->inv_j[c] = -1; tab->inv_slot[c] = 0; }
  for (int j = 0; j < n_off; ++j) {
    const int last = src[j] + (mode[j] == 0 ? 1 : 0);
    if (src[j] < 0 || last >= n_out || (mode[j] == 0 && last >= n_sp) || (mode[j] != 0 && mode[j] != 1))
      return set_error("lsnet_pred_reg: bad table entry %d (src=%d mode=%d)", j, src[j], mode[j]);
    tab->src[j] = static_cast<short>(src[j]); tab->mode[j] = static_cast<signed char>(mode[j]); tab->base[j] = base[j];
    if (mode[j] == 0) {
      tab->inv_j[src[j]] = static_cast<signed char>(j); tab->inv_slot[src[j]] = 0;
      tab->inv_j[src[j] + 1] = static_cast<signed char>(j); tab->inv_slot[src[j] + 1] = 1;
    } else {
      tab->inv_j[src[j]] = static_cast<signed char>(j); tab->inv_slot[src[j]] = 2;
    }
  }
  return 0;
}

extern "C" int lsnet_pred_reg_fwd(const float* o, long long ldo, long long P, int n_sp, int n_out, int n_off,
                                  const int* src, const int* mode, const float* base, float* sp, long long ldsp,
                                  float* off, long long ldoff, void* stream) {
  if (P <= 0) return 0;
  PredRegTab tab;
  if (int rc = fill_tab(&tab, n_off, src, mode, base, n_sp, n_out)) return rc;
  pred_reg_fwd_kernel<<<ew_grid(P * (n_sp + n_off)), 256, 0, static_cast<cudaStream_t>(stream)>>>(o, ldo, P, n_sp, tab, sp,
                                                                                                  ldsp, off, ldoff);
  return check_launch("pred_reg_fwd");
}

extern "C" int lsnet_pred_reg_bwd(const float* o, long long ldo, long long P, int n_sp, int n_out, int n_off,
                                  const int* src, const int* mode, const float* base, const float* gsp, long long ldgsp,
                                  const float* goff, long long ldgoff, float gradient_mul, float* go, long long ldgo,
                                  void* stream) {
  if (P <= 0) return 0;
  PredRegTab tab;
  if (int rc = fill_tab(&tab, n_off, src, mode, base, n_sp, n_out)) return rc;
  pred_reg_bwd_kernel<<<ew_grid(P * n_out), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      o, ldo, P, n_sp, n_out, tab, gsp, ldgsp, goff, ldgoff, gradient_mul, go, ldgo);
  return check_launch("pred_reg_bwd");
}

extern "C" int lsnet_add_softplus(const float* t, long long ldt, const float* s, long long lds, const float* gy,
                                  long long ldgy, long long P, int C, float* out, long long ldout, void* stream) {
  if (P <= 0 || C <= 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (gy) add_softplus_kernel<true><<<ew_grid(P * C), 256, 0, st>>>(t, ldt, s, lds, gy, ldgy, P, C, out, ldout);
  else add_softplus_kernel<false><<<ew_grid(P * C), 256, 0, st>>>(t, ldt, s, lds, nullptr, 0, P, C, out, ldout);
  return check_launch("add_softplus");
}


// ---------------------------------------------------------------------------------------------------------------
// FPN top-down pathway (mmdet/models/necks/fpn.py:180-192): fine += F.interpolate(coarse, size=fine.shape, mode='nearest').
// Pixel-major bf16 maps, 8 channels (16 bytes) per thread.  Source index of PyTorch's nearest mode:
// min(floor(dst * (in / out)), in - 1) with the scale in fp32.  The adjoint sums the fine pixels of every coarse cell.
// ---------------------------------------------------------------------------------------------------------------
namespace lsn {
__device__ __forceinline__ int nearest_src(int dst, float scale, int in_size) {
  const int s = static_cast<int>(floorf(static_cast<float>(dst) * scale));
  return s < in_size - 1 ? s : in_size - 1;
}
__device__ __forceinline__ uint4 add_bf16x8(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {        // fp32 add, one rounding (what torch's bf16 add does)
    const float2 x = __bfloat1622float2(pa[i]), y = __bfloat1622float2(pb[i]);
    pr[i] = __floats2bfloat162_rn(x.x + y.x, x.y + y.y);
  }
  return r;
}

__global__ void upsample_add_kernel(const __nv_bfloat16* __restrict__ fine, long long ldf,
                                    const __nv_bfloat16* __restrict__ coarse, long long ldc, int B, int Hf, int Wf, int Hc,
                                    int Wc, int C, __nv_bfloat16* __restrict__ out, long long ldo) {
  const int vpp = C / 8;
  const long long n = static_cast<long long>(B) * Hf * Wf * vpp;
  const float sh = static_cast<float>(Hc) / static_cast<float>(Hf), sw = static_cast<float>(Wc) / static_cast<float>(Wf);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    long long p = i / vpp;
    const int w = static_cast<int>(p % Wf); p /= Wf;
    const int h = static_cast<int>(p % Hf);
    const int b = static_cast<int>(p / Hf);
    const long long pf = (static_cast<long long>(b) * Hf + h) * Wf + w;
    const long long pc = (static_cast<long long>(b) * Hc + nearest_src(h, sh, Hc)) * Wc + nearest_src(w, sw, Wc);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(fine + pf * ldf) + v);
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(coarse + pc * ldc) + v);
    reinterpret_cast<uint4*>(out + pf * ldo)[v] = add_bf16x8(a, c);
  }
}

// gc[b, s_h, s_w, :] = sum of g over the fine pixels whose nearest source is (s_h, s_w)   (fp32 accumulation)
__global__ void upsample_add_bwd_kernel(const __nv_bfloat16* __restrict__ g, long long ldg, int B, int Hf, int Wf, int Hc,
                                        int Wc, int C, __nv_bfloat16* __restrict__ gc, long long ldgc) {
  const int vpp = C / 8;
  const long long n = static_cast<long long>(B) * Hc * Wc * vpp;
  const float sh = static_cast<float>(Hc) / static_cast<float>(Hf), sw = static_cast<float>(Wc) / static_cast<float>(Wf);
  const int rh = Hf / Hc + 2, rw = Wf / Wc + 2;      // a cell's fine pixels lie within this many rows / columns
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpp);
    long long p = i / vpp;
    const int cw = static_cast<int>(p % Wc); p /= Wc;
    const int ch = static_cast<int>(p % Hc);
    const int b = static_cast<int>(p / Hc);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int h0 = max(0, static_cast<int>(static_cast<float>(ch) / sh) - 1), w0 = max(0, static_cast<int>(static_cast<float>(cw) / sw) - 1);
    for (int h = h0; h < min(Hf, h0 + rh + 2); ++h) {
      if (nearest_src(h, sh, Hc) != ch) continue;
      for (int w = w0; w < min(Wf, w0 + rw + 2); ++w) {
        if (nearest_src(w, sw, Wc) != cw) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(g + ((static_cast<long long>(b) * Hf + h) * Wf + w) * ldg) + v);
        const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(e[k]);
          acc[2 * k] += f.x; acc[2 * k + 1] += f.y;
        }
      }
    }
    reinterpret_cast<uint4*>(gc + ((static_cast<long long>(b) * Hc + ch) * Wc + cw) * ldgc)[v] =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                   pack_bf16x2(acc[6], acc[7]));
  }
}
}  // namespace lsn

extern "C" int lsnet_upsample_add_nhwc_bf16(const void* fine, long long ldf, const void* coarse, long long ldc, int B, int Hf,
                                            int Wf, int Hc, int Wc, int C, void* out, long long ldo, void* stream) {
  if (B <= 0 || Hf <= 0 || Wf <= 0) return 0;
  if ((C % 8) || (ldf % 8) || (ldc % 8) || (ldo % 8) || Hc <= 0 || Wc <= 0)
    return lsn::set_error("lsnet_upsample_add_nhwc_bf16: C %% 8 == 0 and 16-byte aligned pitches required");
  const long long n = static_cast<long long>(B) * Hf * Wf * (C / 8);
  const int grid = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  lsn::upsample_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(fine), ldf, static_cast<const __nv_bfloat16*>(coarse), ldc, B, Hf, Wf, Hc, Wc, C,
      static_cast<__nv_bfloat16*>(out), ldo);
  return lsn::check_launch("upsample_add");
}

extern "C" int lsnet_upsample_add_bwd_nhwc_bf16(const void* g, long long ldg, int B, int Hf, int Wf, int Hc, int Wc, int C,
                                                void* gc, long long ldgc, void* stream) {
  if (B <= 0 || Hc <= 0 || Wc <= 0) return 0;
  if ((C % 8) || (ldg % 8) || (ldgc % 8)) return lsn::set_error("lsnet_upsample_add_bwd_nhwc_bf16: alignment");
  const long long n = static_cast<long long>(B) * Hc * Wc * (C / 8);
  const int grid = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  lsn::upsample_add_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(g), ldg, B, Hf, Wf, Hc, Wc, C, static_cast<__nv_bfloat16*>(gc), ldgc);
  return lsn::check_launch("upsample_add_bwd");
}

// ---- input side: Normalize + Pad + to-tensor of the train pipeline in one pass ----------------------------------------
namespace lsn {
// Thread = 4 consecutive pixels of one row: 12 input bytes (3 aligned 32-bit loads, W % 4 == 0) -> 12 floats (3 float4
// stores); a warp reads 384 and writes 1536 contiguous bytes.  Pixels beyond an image's own (h, w) are written as 0 —
// the pad value of the normalised image — whatever the staging buffer holds there.
__global__ void image_prep_u8_kernel(const uint32_t* __restrict__ src, const int* __restrict__ hw, int B, int H, int W,
                                     double m0, double m1, double m2, double s0, double s1, double s2, int to_rgb,
                                     float4* __restrict__ dst) {
  const int qpr = W / 4;
  const long long n = static_cast<long long>(B) * H * qpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(i % qpr);
    long long r = i / qpr;
    const int y = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    const int h = __ldg(hw + 2 * b), w = __ldg(hw + 2 * b + 1);
    float o[12];
    if (y < h && 4 * q < w) {
      const uint32_t a0 = __ldg(src + 3 * i), a1 = __ldg(src + 3 * i + 1), a2 = __ldg(src + 3 * i + 2);
      float v[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = static_cast<float>((a0 >> (8 * k)) & 0xffu);
        v[4 + k] = static_cast<float>((a1 >> (8 * k)) & 0xffu);
        v[8 + k] = static_cast<float>((a2 >> (8 * k)) & 0xffu);
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int px = k / 3, ch = k % 3;                       // output channel ch of pixel px
        const float x = to_rgb ? v[px * 3 + 2 - ch] : v[px * 3 + ch];
        const double m = ch == 0 ? m0 : (ch == 1 ? m1 : m2), s = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
        // double arithmetic, one rounding to float: bit-identical to cv2.subtract / cv2.multiply with float64 scalars
        o[k] = (4 * q + px < w) ? static_cast<float>(__dmul_rn(__dsub_rn(static_cast<double>(x), m), s)) : 0.f;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 12; ++k) o[k] = 0.f;
    }
    float4* d = dst + 3 * i;
    d[0] = make_float4(o[0], o[1], o[2], o[3]);
    d[1] = make_float4(o[4], o[5], o[6], o[7]);
    d[2] = make_float4(o[8], o[9], o[10], o[11]);
  }
}
}  // namespace lsn

extern "C" int lsnet_image_prep_u8(const void* src_u8, const int* hw, int B, int H, int W, double mean0, double mean1,
                                   double mean2, double stdinv0, double stdinv1, double stdinv2, int to_rgb,
                                   void* dst_f32, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  if (W % 4) return lsn::set_error("lsnet_image_prep_u8: W %% 4 == 0 required (the canvas is padded to a multiple of 32)");
  if ((reinterpret_cast<uintptr_t>(src_u8) & 3) || (reinterpret_cast<uintptr_t>(dst_f32) & 15))
    return lsn::set_error("lsnet_image_prep_u8: 4-byte aligned source / 16-byte aligned destination required");
  const long long n = static_cast<long long>(B) * H * (W / 4);
  const int grid = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
  lsn::image_prep_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint32_t*>(src_u8), hw, B, H, W, mean0, mean1, mean2, stdinv0, stdinv1, stdinv2, to_rgb,
      static_cast<float4*>(dst_f32));
  return lsn::check_launch("image_prep_u8");
}
