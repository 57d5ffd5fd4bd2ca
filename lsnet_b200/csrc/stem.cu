// ResNet stem on tcgen05 (sm_100a): conv 7x7 / stride 2 / pad 3, 3 -> 64 channels, with the frozen BatchNorm folded in
// (scale in the weights, shift as bias) and ReLU in the epilogue, plus the 3x3 / stride-2 max-pool that follows it
// (mmdet/models/backbones/resnet.py:509-520, 619-623: conv1 -> norm1 -> relu -> maxpool).
//
// The input has 3 channels, so no TMA box can form a K-major operand tile; the kernel is an implicit GEMM whose A
// operand is BUILT in shared memory by producer warps (the same producer / fence.proxy.async / single-thread
// tcgen05.mma structure as the fused deformable convolution):
//   CTA tile   = 8 x 16 conv outputs (UMMA M = 128), persistent over tiles
//   patch      = the 21 x 37 x 3 input window of the tile, read once from the fp32 / bf16 image (any strides: NCHW or
//                NHWC), converted to bf16 and laid out [row][col*3 + ch] in shared memory (the next tile's window is
//                prefetched into registers while the current A tile is built)
//   K layout   = ky * 24 + (kx*3 + ch): the 21 values of one filter row are CONTIGUOUS in the patch
//                (patch[2r+ky][(2c)*3 ...]), so an A row is seven 48-byte copies; the 3 pad entries per filter row and
//                K 168..191 meet zero weights.  K = 192 = three SWIZZLE_128B K blocks.
//   B operand  = folded weights [64, 192] bf16, loaded once per CTA by TMA and kept resident
//   accumulator= 128 x 64 fp32 in TMEM, two stages; epilogue warps add the shift, apply ReLU and store bf16 NHWC rows
// Warp roles (416 threads): warp 0 TMA + TMEM alloc + MMA issuer, warps 1-4 epilogue, warps 5-12 producers.
#include "common.cuh"
#include "lsnet_internal.h"

namespace lsn {

constexpr int S_TH = 8, S_TW = 16;                 // conv outputs per tile
constexpr int S_PH = 2 * S_TH + 5, S_PW = 2 * S_TW + 5;   // 21 x 37 input window
constexpr int S_PITCH = 240;                       // bytes per patch row (111 bf16 used; reads run to byte 228)
constexpr int S_KBLK = 3;                          // 192 / 64
constexpr int S_ASTAGE = S_KBLK * 128 * 128;       // 48 KB: three [128 rows x 128 B] K blocks
constexpr int S_WBYTES = S_KBLK * 64 * 128;        // 24 KB
constexpr int S_PROD = 256;                        // producer threads
constexpr int S_THREADS = 32 + 128 + S_PROD;
constexpr int S_SMEM = 2 * S_ASTAGE + S_WBYTES + S_PH * S_PITCH + 1024 + 256;
constexpr int S_PLOADS = (S_PH * S_PW * 3 + S_PROD - 1) / S_PROD;   // window elements per producer thread (10)

struct StemArgs {
  const void* x;
  long long sb, sc, sh, sw;      // input strides in elements
  int x_bf16;
  int B, H, W, Ho, Wo, tiles_h, tiles_w, num_tiles;
  const float* bias;             // [64]
  __nv_bfloat16* out;            // [B, Ho, Wo, 64]
};

__device__ __forceinline__ float stem_load(const StemArgs& p, int b, int ch, int h, int w) {
  if (h < 0 || h >= p.H || w < 0 || w >= p.W) return 0.f;
  const long long off = b * p.sb + ch * p.sc + h * p.sh + w * p.sw;
  return p.x_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.x)[off])
                  : __ldg(reinterpret_cast<const float*>(p.x) + off);
}

__global__ void __launch_bounds__(S_THREADS, 1)
stem_conv_kernel(const __grid_constant__ CUtensorMap tmW, const StemArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem + 2 * S_ASTAGE;
  uint8_t* patch = sW + S_WBYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(patch + S_PH * S_PITCH);   // [2] A stage built
  uint64_t* empty_bar = full_bar + 2;                                          // [2] A stage consumed
  uint64_t* tfull_bar = empty_bar + 2;                                         // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;                                        // [2] accumulator drained
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], S_PROD / 32);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // zero both A stages (the pad chunks 21..23 of every row are never written again) and the patch tail
  for (int i = threadIdx.x; i < (2 * S_ASTAGE) / 16; i += S_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < (S_PH * S_PITCH) / 16; i += S_THREADS)
    reinterpret_cast<uint4*>(patch)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_bar, S_WBYTES);
      for (int kb = 0; kb < S_KBLK; ++kb) tma_load_2d(sW + kb * (64 * 128), &tmW, w_bar, kb * 64, 0);
      mbar_wait(w_bar, 0);
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        mbar_wait(&tempty_bar[s], ph ^ 1);
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + s * S_ASTAGE), sB = smem_u32(sW);
#pragma unroll
        for (int kb = 0; kb < S_KBLK; ++kb) {
          const uint64_t adesc = umma_desc_sw128(sA + kb * (128 * 128), 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sB + kb * (64 * 128), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + s * 64, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
        umma_commit(&tfull_bar[s]);
      }
    }
  } else if (warp < 5) {
    // ===================== epilogue: TMEM -> (+shift, ReLU) -> bf16 NHWC =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                  // tile pixel = accumulator row
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      const int b = tile / per_img, t2 = tile % per_img;
      const int h = (t2 / p.tiles_w) * S_TH + r / S_TW, w = (t2 % p.tiles_w) * S_TW + r % S_TW;
      mbar_wait(&tfull_bar[s], ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(s * 64);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        if (h < p.Ho && w < p.Wo) {
          uint4* o = reinterpret_cast<uint4*>(p.out + ((static_cast<long long>(b) * p.Ho + h) * p.Wo + w) * 64 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(__uint_as_float(v[8 * j + i]) + __ldg(p.bias + c0 + 8 * j + i), 0.f);
            o[j] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[s]);
    }
  } else {
    // ===================== producers: input window -> bf16 patch -> swizzled A tile =====================
    const int t = threadIdx.x - 160;
    const uint32_t patch_s = smem_u32(patch);
    float reg[S_PLOADS];
    auto fetch = [&](int tile) {
      const int b = tile / per_img, t2 = tile % per_img;
      const int h0 = (t2 / p.tiles_w) * S_TH * 2 - 3, w0 = (t2 % p.tiles_w) * S_TW * 2 - 3;
#pragma unroll
      for (int i = 0; i < S_PLOADS; ++i) {
        const int e = t + i * S_PROD;                 // e = (ch * PH + row) * PW + col  (col fastest: coalesced along W)
        float v = 0.f;
        if (e < S_PH * S_PW * 3) {
          const int col = e % S_PW, rr = e / S_PW;
          v = stem_load(p, b, rr / S_PH, h0 + rr % S_PH, w0 + col);
        }
        reg[i] = v;
      }
    };
    if (blockIdx.x < p.num_tiles) fetch(blockIdx.x);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      // registers -> patch (every producer is done reading the previous patch: barrier at the end of the last build)
#pragma unroll
      for (int i = 0; i < S_PLOADS; ++i) {
        const int e = t + i * S_PROD;
        if (e < S_PH * S_PW * 3) {
          const int col = e % S_PW, rr = e / S_PW;
          const int ch = rr / S_PH, row = rr % S_PH;
          reinterpret_cast<__nv_bfloat16*>(patch + row * S_PITCH)[col * 3 + ch] = __float2bfloat16(reg[i]);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(S_PROD) : "memory");
      const int next = tile + gridDim.x;
      if (next < p.num_tiles) fetch(next);            // in flight while the A tile is built
      mbar_wait(&empty_bar[s], ph ^ 1);               // the tensor core is done with this A stage
      const uint32_t sA = smem_u32(smem + s * S_ASTAGE);
      for (int u = t; u < 128 * 7; u += S_PROD) {
        const int pix = u & 127, ky = u >> 7;
        const int pr = pix / S_TW, pc = pix % S_TW;
        const uint32_t src = patch_s + (2 * pr + ky) * S_PITCH + 12 * pc;
#pragma unroll
        for (int sub = 0; sub < 3; ++sub) {
          const uint4 v = make_uint4(lds32(src + 16 * sub), lds32(src + 16 * sub + 4), lds32(src + 16 * sub + 8),
                                     lds32(src + 16 * sub + 12));
          const int qk = ky * 3 + sub;                // 16-byte chunk index along K (0..20)
          sts128(sA + (qk >> 3) * (128 * 128) + pix * 128 + (((qk & 7) ^ (pix & 7)) << 4), v);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
      asm volatile("bar.sync 1, %0;" ::"n"(S_PROD) : "memory");      // patch free for the next window
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// 3x3 / stride 2 / pad 1 max-pool over NHWC bf16, 8 channels (16 bytes) per thread
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C,
                                    __nv_bfloat16* __restrict__ y, int Ho, int Wo) {
  const int vpp = C / 8;
  const long long n = static_cast<long long>(B) * Ho * Wo * vpp;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = static_cast<int>(i % vpp);
  long long pix = i / vpp;
  const int wo = static_cast<int>(pix % Wo); pix /= Wo;
  const int ho = static_cast<int>(pix % Ho);
  const int b = static_cast<int>(pix / Ho);
  __nv_bfloat162 m[4];
  const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
  for (int k = 0; k < 4; ++k) m[k] = ninf;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int h = 2 * ho - 1 + dy;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int w = 2 * wo - 1 + dx;
      if (w < 0 || w >= W) continue;
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(b) * H + h) * W + w) * C) + v);
      const __nv_bfloat162* e = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int k = 0; k < 4; ++k) m[k] = __hmax2(m[k], e[k]);
    }
  }
  uint4 o;
  __nv_bfloat162* oe = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int k = 0; k < 4; ++k) oe[k] = m[k];
  reinterpret_cast<uint4*>(y + ((static_cast<long long>(b) * Ho + ho) * Wo + wo) * C)[v] = o;
}

}  // namespace lsn

using namespace lsn;

// x: [B, 3, H, W] image, fp32 or bf16, element strides (sb, sc, sh, sw); Wp: bf16 [64, 192], column ky*24 + kx*3 + ch
// (zero elsewhere) = folded weight; bias fp32 [64]; out bf16 NHWC [B, Ho, Wo, 64], Ho = (H + 6 - 7)/2 + 1.
extern "C" int lsnet_stem_conv7x7s2_bf16(const void* x, int x_bf16, long long sb, long long sc, long long sh, long long sw,
                                         int B, int H, int W, const void* Wp, const float* bias, void* out, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(stem_conv): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  StemArgs a{};
  a.x = x; a.sb = sb; a.sc = sc; a.sh = sh; a.sw = sw; a.x_bf16 = x_bf16;
  a.B = B; a.H = H; a.W = W;
  a.Ho = (H - 1) / 2 + 1; a.Wo = (W - 1) / 2 + 1;
  a.tiles_h = (a.Ho + S_TH - 1) / S_TH; a.tiles_w = (a.Wo + S_TW - 1) / S_TW;
  a.num_tiles = B * a.tiles_h * a.tiles_w;
  a.bias = bias;
  a.out = static_cast<__nv_bfloat16*>(out);
  CUtensorMap tmW;
  if (int rc = make_map_2d(&tmW, Wp, 64, 192, 192, 64, 64)) return rc;
  const int grid = a.num_tiles < num_sms() ? a.num_tiles : num_sms();
  stem_conv_kernel<<<grid, S_THREADS, S_SMEM, static_cast<cudaStream_t>(stream)>>>(tmW, a);
  return check_launch("stem_conv");
}

extern "C" int lsnet_maxpool3x3s2_nhwc_bf16(const void* x, int B, int H, int W, int C, void* y, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  if (C % 8) return set_error("lsnet_maxpool3x3s2_nhwc_bf16: C %% 8 != 0");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long n = static_cast<long long>(B) * Ho * Wo * (C / 8);
  maxpool3x3s2_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), B, H, W, C, static_cast<__nv_bfloat16*>(y), Ho, Wo);
  return check_launch("maxpool3x3s2");
}
