// Fused deformable convolution, forward (sm_100a): the bilinear-offset gather IS the A-operand producer of a tcgen05
// GEMM.  Gather warps sample x at the learned positions, interpolate on packed fp32x2 FMAs and write bf16 rows straight
// into the SWIZZLE_128B K-major shared-memory tile that `tcgen05.mma` reads -- the [pixels, taps*C] column matrix of the
// reference (deform_conv_cuda.cpp:655-684: im2col kernel -> HBM columns -> addmm_) never exists in HBM.
//
//   out[p, n] = sum_{tap, c} ( sum_q w_q(p,tap) * x[corner_q(p,tap), c] ) * Wp[n, tap*C + c]  (+ bias[n]) (ReLU)
//
// One CTA = one 128-pixel patch (TH x TW) of one image, persistent over patches.  Warp roles (704 threads):
//   warp 0      TMA producer of the weight tiles (B operand, [BN x 64] bf16 boxes, SWIZZLE_128B)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, fp32 accumulators in TMEM)
//   warps 2-5   epilogue: tcgen05.ld -> (+bias, ReLU) -> bf16/fp32 rows of the NHWC output (or a channel slice of it)
//   warps 6-21  gather: per patch the sampling geometry of every (pixel, tap) is derived ONCE into shared-memory records
//               (4 clamped corner pixel indices + 4 mask-folded bilinear weights); the K loop then runs channel-block
//               outer / tap inner, so that the 128-byte channel slices of the ~(TH+2)x(TW+2) cells a patch touches stay
//               L1-resident across the 9 taps.  A lane owns 8 channels (16 bytes) of one pixel: 4 independent 16-byte
//               corner loads per (pixel, tap), 8 in flight per lane; eight lanes cover the 64-channel K block of a pixel,
//               a warp instruction four pixels.
// Full barrier of a stage = 16 gather-warp arrivals + the TMA transaction bytes; generic-proxy tile writes are made
// visible to the tensor core with fence.proxy.async before the arrive.  Two TMEM accumulator stages overlap the epilogue
// of patch i with the MMAs of patch i+1.
//
// Semantics: DCNv1 (mask NULL), DCNv2 (mask / mask logits), LSNet pyramid DCN (scale_h, scale_w; input extent (H,W)
// decoupled from the sampling grid (Ho,Wo)) -- arithmetic of deform_conv_cuda_kernel.cu:190-297, 847-910 (sampling) and
// deform_conv_cuda.cpp:673-691, 890-906 (GEMM + bias).
#include <stdlib.h>

#include "common.cuh"
#include "dcn_common.cuh"
#include "lsnet_internal.h"

namespace lsn {

constexpr int FM = 128;             // output pixels per tile = UMMA M
constexpr int FK = 64;              // K block: 64 bf16 channels of one tap = one SWIZZLE_128B row
constexpr int F_MAXTAPS = 9;
constexpr int F_GW = 16;            // gather warps
constexpr int F_EPI0 = 2;           // first epilogue warp (warps 2..5 -> TMEM lane quarters 2,3,0,1)
constexpr int F_G0 = 6;             // first gather warp
constexpr int F_THREADS = (F_G0 + F_GW) * 32;
constexpr int F_UPW = (FM / 4) / F_GW;   // (4-pixel units per K block) / gather warps = units per warp per K block
static_assert(F_UPW * F_GW * 4 == FM, "gather warps must tile the 128 pixels");

struct FusedFwdArgs {
  DcnGeom g;
  const __nv_bfloat16* x;
  const float* offset;
  const float* mask;
  int N;                              // logical output channels
  int tiles_h, tiles_w, TH, TW, tw_shift, num_tiles;   // TW = 1 << tw_shift
  int taps, cblks;
  // grouped (block-diagonal weights on 64-channel blocks): every channel block is its own accumulation group -- 9 K
  // blocks -> BN = 64 output columns at column offset cb * 64; weights packed [C, taps * 64].  The blocks are independent,
  // so a patch's blocks are dealt to cb_parts work items (small maps: a patch alone would run C/64 * 9 K blocks serially)
  int grouped, cb_parts, cb_per_part;
  void* out;
  long long ldc;
  int out_fp32, relu;
  const float* bias;
  __nv_bfloat16* col;                 // optional side output: the bf16 column matrix [pixels, taps*C] (or null)
  double* gn_sums;                    // optional: GroupNorm statistics of the output (before ReLU), [B, G, 2] fp64, pre-zeroed
  int gn_cpg, gn_G;
};

template <int BN, int STAGES>
struct FCfg {
  static constexpr int kABytes = FM * FK * 2;
  static constexpr int kBBytes = BN * FK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kStagingBytes = 4 * 32 * 80;                    // per epilogue warp: 32 rows x (64 B + 16 B pad)
  static constexpr int kGeoBytes = F_MAXTAPS * FM * 32 + FM * 4;       // records + linear pixel index per tile row
  static constexpr int kSmemBytes = STAGES * kStageBytes + kStagingBytes + kGeoBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN, int STAGES, bool SAVE_COL>
__global__ void __launch_bounds__(F_THREADS, 1)
dcn_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmB, const FusedFwdArgs p) {
  using Cfg = FCfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + STAGES * Cfg::kStageBytes);
  uint4* gidx = reinterpret_cast<uint4*>(smem + STAGES * Cfg::kStageBytes + Cfg::kStagingBytes);   // [tap][pix]
  float4* gwt = reinterpret_cast<float4*>(gidx + F_MAXTAPS * FM);                                  // [tap][pix]
  int* gpix = reinterpret_cast<int*>(gwt + F_MAXTAPS * FM);                                        // [pix]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(gpix + FM);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], F_GW + 1);      // 16 gather warps + the TMA producer's expect_tx arrive
      mbar_init(&empty_bar[s], 1);            // tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);           // one elected lane per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_k_iters = p.taps * p.cblks;
  const int k_per_group = p.grouped ? p.taps : num_k_iters;     // K blocks per accumulator
  const int per_img = p.tiles_h * p.tiles_w;
  const int num_items = p.num_tiles * p.cb_parts;               // work item = (patch, range of channel blocks)
  // item -> patch index, first / one-past-last channel block
  auto item_of = [&](int item, int* tile, int* cb0, int* cb1) {
    *tile = item / p.cb_parts;
    const int part = item - *tile * p.cb_parts;
    *cb0 = part * p.cb_per_part;
    *cb1 = min(p.cblks, *cb0 + p.cb_per_part);
  };

  if (warp == 0) {
    // ===================== TMA producer: weight tiles =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tile, cb0, cb1;
        item_of(item, &tile, &cb0, &cb1);
        for (int cb = cb0; cb < cb1; ++cb) {
          for (int tap = 0; tap < p.taps; ++tap) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sB = smem + stage * Cfg::kStageBytes + Cfg::kABytes;
            mbar_expect_tx(&full_bar[stage], Cfg::kBBytes);
            if (p.grouped) tma_load_2d(sB, &tmB, &full_bar[stage], tap * FK, cb * FK);
            else tma_load_2d(sB, &tmB, &full_bar[stage], tap * p.g.C + cb * FK, 0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(FM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int tile, cb0, cb1;
        item_of(item, &tile, &cb0, &cb1);
        const int acc_groups = p.grouped ? cb1 - cb0 : 1;
        for (int grp = 0; grp < acc_groups; ++grp) {
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
          for (int it = 0; it < k_per_group; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sA = smem_u32(smem + stage * Cfg::kStageBytes);
            const uint32_t sB = sA + Cfg::kABytes;
            const uint64_t adesc = umma_desc_sw128(sA, 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(sB, 16, 1024);
#pragma unroll
            for (int k = 0; k < FK / 16; ++k)
              umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (it == k_per_group - 1) umma_commit(&tfull_bar[as]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++as == 2) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else if (warp < F_G0) {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;     // accumulator row = pixel of the patch handled by this thread
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t stg_s = smem_u32(staging + (warp - F_EPI0) * (32 * 20));      // explicit ld/st.shared (not generic LD.E/ST.E)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tile, cb0, cb1;
      item_of(item, &tile, &cb0, &cb1);
      const int acc_groups = p.grouped ? cb1 - cb0 : 1;
      const int b = tile / per_img, t2 = tile % per_img;
      const int h_base = (t2 / p.tiles_w) * p.TH, w_base = (t2 % p.tiles_w) * p.TW;
      auto row_of = [&](int rr, bool* ok) -> long long {
        const int h = h_base + (rr >> p.tw_shift), w = w_base + (rr & (p.TW - 1));
        *ok = (h < p.g.Ho) && (w < p.g.Wo);
        return (static_cast<long long>(b) * p.g.Ho + h) * p.g.Wo + w;
      };
      bool valid;
      const long long row = row_of(r, &valid);
      for (int grp = 0; grp < acc_groups; ++grp) {
        const int n_base = p.grouped ? (cb0 + grp) * BN : 0;   // first output column (grouped: the channel block)
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          const int c = n_base + c0;
          if (c >= p.N) continue;     // warp-uniform
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < p.N) f[j] += __ldg(p.bias + c + j);
          }
          if (p.gn_sums) gn_epilogue_sums(f, valid, b, b, b, c, p.N, p.gn_cpg, p.gn_G, p.gn_sums, lane);
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.out_fp32) {
            if (valid) {
              float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + c;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (c + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            }
          } else {
            // stage the 32x32 bf16 block in shared memory (row pitch 80 B), then store 8 rows x 64 contiguous bytes per
            // warp instruction instead of 32 rows x 16 bytes
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(stg_s + (lane * 20 + j * 4) * 4,
                     make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7])));
            __syncwarp();
            const int seg = lane & 3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = (lane >> 2) + 8 * i;
              bool ok;
              const long long grow = row_of(q * 32 + rr, &ok);
              if (ok && c + seg * 8 < p.N) {
                const uint4 val = lds128(stg_s + (rr * 20 + seg * 4) * 4);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + grow * p.ldc + c + seg * 8) = val;
              }
            }
            __syncwarp();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== gather warps (6..21): the A-operand producer =====================
    const int gwarp = warp - F_G0;
    const int gtid = threadIdx.x - F_G0 * 32;
    const int grp = lane >> 3, sub = lane & 7;
    const DcnGeom& g = p.g;
    const uint32_t ldxb = static_cast<uint32_t>(g.ldx) * 2u;           // pixel pitch in bytes (host-checked < 2^31)
    const char* xlane = reinterpret_cast<const char*>(p.x) + sub * 16;
    const uint32_t smem_s = smem_u32(smem), gidx_s = smem_u32(gidx), gwt_s = smem_u32(gwt), gpix_s = smem_u32(gpix);
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      int tile, cb0, cb1;
      item_of(item, &tile, &cb0, &cb1);
      const int item_k_iters = (cb1 - cb0) * p.taps;
      const int b = tile / per_img, t2 = tile % per_img;
      const int h_base = (t2 / p.tiles_w) * p.TH, w_base = (t2 % p.tiles_w) * p.TW;
      // every gather warp is done with the previous patch's records
      named_bar_sync(1, F_GW * 32);
      for (int e = gtid; e < p.taps * FM; e += F_GW * 32) {
        const int tap = e / FM, pix = e - tap * FM;
        const int ho = h_base + (pix >> p.tw_shift), wo = w_base + (pix & (p.TW - 1));
        uint4 ri = make_uint4(0u, 0u, 0u, 0u);
        float4 rw = make_float4(0.f, 0.f, 0.f, 0.f);
        int lin = -1;
        if (ho < g.Ho && wo < g.Wo) {
          const long long pl = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
          lin = static_cast<int>(pl);
          float h, w;
          sample_pos(g, p.offset + pl * g.ldo, 0, tap, ho, wo, &h, &w);
          uint32_t pi[4];
          float wt[4];
          corner_pixels(g, b, h, w, pi, wt);
          float m = 1.f;
          if (p.mask) m = load_mask(g, p.mask + pl * g.ldm + tap);
          rw = make_float4(wt[0] * m, wt[1] * m, wt[2] * m, wt[3] * m);
          ri = make_uint4(pi[0], pi[1], pi[2], pi[3]);
        }
        gidx[e] = ri;
        gwt[e] = rw;
        if (tap == 0) gpix[pix] = lin;
      }
      named_bar_sync(1, F_GW * 32);

      // Software pipeline over the K blocks (channel-block outer, tap inner): the corner loads of K block i+1 are issued
      // into a unit's registers right after K block i's values have been consumed from them, so every lane always has
      // 4 * F_UPW independent 16-byte loads in flight while it interpolates.
      uint4 ld[F_UPW][4];
      float4 wq[F_UPW];
      const unsigned long long x0 = reinterpret_cast<unsigned long long>(xlane);
#define LSN_ISSUE(u, tap_, cb_)                                                             \
  {                                                                                         \
    const int pix_ = (gwarp * F_UPW + (u)) * 4 + grp;                                       \
    const uint4 ci_ = lds128(gidx_s + ((tap_) * FM + pix_) * 16);                           \
    wq[u] = lds128f(gwt_s + ((tap_) * FM + pix_) * 16);                                     \
    const unsigned long long xc_ = x0 + (cb_) * (FK * 2);                                   \
    ld[u][0] = ldg128_at(xc_, ci_.x, ldxb);                                                 \
    ld[u][1] = ldg128_at(xc_, ci_.y, ldxb);                                                 \
    ld[u][2] = ldg128_at(xc_, ci_.z, ldxb);                                                 \
    ld[u][3] = ldg128_at(xc_, ci_.w, ldxb);                                                 \
  }
#pragma unroll
      for (int u = 0; u < F_UPW; ++u) LSN_ISSUE(u, 0, cb0)
      int tap = 0, cb = cb0;
      for (int it = 0; it < item_k_iters; ++it) {
        int ntap = tap + 1, ncb = cb;
        if (ntap == p.taps) { ntap = 0; ++ncb; }
        const bool has_next = it + 1 < item_k_iters;
        mbar_wait(&empty_bar[stage], phase ^ 1);      // the tensor core is done reading this stage
        const uint32_t sA = smem_s + stage * Cfg::kStageBytes;
#pragma unroll
        for (int u = 0; u < F_UPW; ++u) {
          const int pix = (gwarp * F_UPW + u) * 4 + grp;
          uint32_t outw[4];
          const float2 w0 = make_float2(wq[u].x, wq[u].x), w1 = make_float2(wq[u].y, wq[u].y);
          const float2 w2 = make_float2(wq[u].z, wq[u].z), w3 = make_float2(wq[u].w, wq[u].w);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t a0 = reinterpret_cast<const uint32_t*>(&ld[u][0])[i];
            const uint32_t a1 = reinterpret_cast<const uint32_t*>(&ld[u][1])[i];
            const uint32_t a2 = reinterpret_cast<const uint32_t*>(&ld[u][2])[i];
            const uint32_t a3 = reinterpret_cast<const uint32_t*>(&ld[u][3])[i];
            // same association as the unfused gather: ((w0*f0 + w1*f1) + w2*f2) + w3*f3, one fused FMA per step
            float2 acc = __fmul2_rn(w0, bf16x2_f2(a0));
            acc = __ffma2_rn(w1, bf16x2_f2(a1), acc);
            acc = __ffma2_rn(w2, bf16x2_f2(a2), acc);
            acc = __ffma2_rn(w3, bf16x2_f2(a3), acc);
            outw[i] = pack_bf16x2(acc.x, acc.y);
          }
          const uint4 val = make_uint4(outw[0], outw[1], outw[2], outw[3]);
          if (has_next) LSN_ISSUE(u, ntap, ncb)
          // SWIZZLE_128B K-major: row = pixel (128 B), 16-byte chunk index XOR (row & 7)
          sts128(sA + pix * 128 + ((sub ^ (pix & 7)) << 4), val);
          if (SAVE_COL) {
            const int lin = static_cast<int>(lds32(gpix_s + pix * 4));
            if (lin >= 0)
              st_stream(p.col + static_cast<long long>(lin) * g.ldcol + tap * g.C + cb * FK + sub * 8, val);
          }
        }
        fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        tap = ntap;
        cb = ncb;
      }
#undef LSN_ISSUE
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =====================================================================================================================
// Fused weight gradient:  dW[n, tap*C + c] += sum_p dY[p, n] * col[p, tap*C + c]  with the columns RE-SAMPLED on the fly
// (the reference re-runs its im2col kernel into HBM for this, deform_conv_cuda.cpp:770-787, 1098-1124).
// MN-major tcgen05 GEMM over pixels: per 64-pixel patch (TH x TW) the A operand dY[64 px, 256 couts] arrives by TMA
// (4-D NHWC map, out-of-range pixels zero-filled, so ragged patches need no masking), the B operand
// col[64 px, 256 channels of ONE tap] is written by the gather warps as four [64 x 64] SWIZZLE_128B sub-tiles.
// Work item = (tap, channel tile of 256, cout group of 256, pixel split); the 256 x 256 fp32 accumulator fills the 512
// TMEM columns (two 128-row halves that share every B tile).  Epilogue: fp32 red.global.add into dW, or -- deterministic
// mode -- plain stores of the per-split partial into a workspace that dcn_wgrad_reduce sums in a fixed order.
// Gather warp w owns pixels 4w..4w+3 of the patch (lane>>3) for all four channel blocks: its lanes derive the sampling
// geometry themselves (no shared records, no block barrier), offsets / mask of the next patch are prefetched.
// =====================================================================================================================
constexpr int WK = 64;                 // pixels per K chunk
constexpr int W_STAGES = 2;
constexpr int W_ABYTES = WK * 256 * 2; // dY  [64 px x 256 couts]
constexpr int W_BBYTES = WK * 256 * 2; // col [64 px x 256 channels]
constexpr int W_STAGE = W_ABYTES + W_BBYTES;
constexpr int W_SMEM = W_STAGES * W_STAGE + 1024 + 256;
static_assert(F_GW * 4 == WK, "one gather warp per 4 pixels of the chunk");

struct FusedWgradArgs {
  DcnGeom g;
  const __nv_bfloat16* x;
  const float* offset;
  const float* mask;
  int M;                               // couts (rows of dW)
  int tiles_h, tiles_w, TH, TW, tw_shift, k_chunks;   // 64-pixel patches: k_chunks = B * tiles_h * tiles_w
  int taps, n_tiles, m_groups, splits, chunks_per_split;
  int diag;                            // grouped weights: only the tiles on the block diagonal (cout group == channel
                                       // tile) are computed, written compactly as out[m, tap * 256 + c_in_tile]
  float* out;                          // dW [M, ldw]  (or the partial workspace in deterministic mode)
  long long ldw;
  long long split_stride;              // 0: accumulate with reds into `out`; else elements between per-split partials
};

__global__ void __launch_bounds__(F_THREADS, 1)
dcn_fused_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const FusedWgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + W_STAGES * W_STAGE);
  uint64_t* empty_bar = full_bar + W_STAGES;
  uint64_t* tfull_bar = empty_bar + W_STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(&full_bar[s], F_GW + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int items = p.taps * p.n_tiles * p.m_groups * p.splits;
  const int per_img = p.tiles_h * p.tiles_w;
  // item -> (split, m_group, n_tile, tap); consecutive CTAs take different splits of the same tap
  auto decode = [&](int item, int* split, int* mg, int* nt, int* tap) {
    *split = item % p.splits; item /= p.splits;
    *mg = item % p.m_groups; item /= p.m_groups;
    *nt = item % p.n_tiles; item /= p.n_tiles;
    *tap = item;
    if (p.diag) *mg = *nt;
  };

  if (warp == 0) {
    // ===================== TMA producer: dY tiles =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int split, mg, nt, tap;
        decode(item, &split, &mg, &nt, &tap);
        const int c_begin = split * p.chunks_per_split;
        const int c_end = min(p.k_chunks, c_begin + p.chunks_per_split);
        for (int ch = c_begin; ch < c_end; ++ch) {
          const int b = ch / per_img, rr = ch % per_img;
          const int h0 = (rr / p.tiles_w) * p.TH, w0 = (rr % p.tiles_w) * p.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * W_STAGE;
          mbar_expect_tx(&full_bar[stage], W_ABYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_4d(sA + j * (WK * 128), &tmA, &full_bar[stage], mg * 256 + j * 64, w0, h0, b);
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, 256, 1, 1);
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int split, mg, nt, tap;
        decode(item, &split, &mg, &nt, &tap);
        const int c_begin = split * p.chunks_per_split;
        const int c_end = min(p.k_chunks, c_begin + p.chunks_per_split);
        if (c_end <= c_begin) continue;
        mbar_wait(tempty_bar, aphase ^ 1);
        tc_fence_after();
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * W_STAGE);
          const uint32_t sB = sA + W_ABYTES;
          // MN-major: LBO = distance between 64-wide M/N chunks (one [64 x 128 B] sub-tile), SBO = 8 K rows
          const uint64_t bdesc = umma_desc_sw128(sB, WK * 128, 1024);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint64_t adesc = umma_desc_sw128(sA + mt * 2 * (WK * 128), WK * 128, 1024);
#pragma unroll
            for (int k = 0; k < WK / 16; ++k)
              umma_bf16(tmem_base + mt * 256, adesc + 128 * k, bdesc + 128 * k, idesc, ((ch - c_begin) | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (ch == c_end - 1) umma_commit(tfull_bar);
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
        aphase ^= 1;
      }
    }
  } else if (warp < F_G0) {
    // ===================== epilogue (warps 2..5): TMEM -> dW =====================
    const int q = warp & 3;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int split, mg, nt, tap;
      decode(item, &split, &mg, &nt, &tap);
      const int c_begin = split * p.chunks_per_split;
      const int c_end = min(p.k_chunks, c_begin + p.chunks_per_split);
      if (c_end <= c_begin) continue;
      mbar_wait(tfull_bar, aphase);
      tc_fence_after();
      float* obase = p.out + static_cast<long long>(split) * p.split_stride;
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int m = mg * 256 + mt * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mt * 256);
#pragma unroll 1
        for (int c = 0; c < 256; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
          const int col0 = nt * 256 + c;
          if (m < p.M && col0 < p.g.C) {
            float* o = obase + static_cast<long long>(m) * p.ldw +
                       (p.diag ? static_cast<long long>(tap) * 256 + c : static_cast<long long>(tap) * p.g.C + col0);
            if (p.split_stride) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(__uint_as_float(v[j])),
                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                             "f"(__uint_as_float(v[j + 3]))
                             : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      aphase ^= 1;
    }
  } else {
    // ===================== gather warps: col[64 px, 256 ch] of one tap =====================
    const int gwarp = warp - F_G0;
    const int grp = lane >> 3, sub = lane & 7;
    const int r = gwarp * 4 + grp;                         // pixel row of the patch handled by this lane group
    const DcnGeom& g = p.g;
    const uint32_t ldxb = static_cast<uint32_t>(g.ldx) * 2u;
    const char* xlane = reinterpret_cast<const char*>(p.x) + sub * 16;
    const uint32_t smem_s = smem_u32(smem);
    const uint32_t row_off = r * 128 + ((sub ^ (r & 7)) << 4);
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int split, mg, nt, tap;
      decode(item, &split, &mg, &nt, &tap);
      const int c_begin = split * p.chunks_per_split;
      const int c_end = min(p.k_chunks, c_begin + p.chunks_per_split);
      if (c_end <= c_begin) continue;
      const unsigned long long xt = reinterpret_cast<unsigned long long>(xlane + nt * 512);
      const int ti = tap / g.kw, tj = tap % g.kw;
      // Pipeline: offsets / mask of patch i+2 are fetched and the geometry of patch i+1 is derived while patch i is
      // gathered; two 4-corner load sets are always in flight per lane (channel blocks rotate through ld[0], ld[1]).
      float oh = 0.f, ow = 0.f, mraw = 0.f;
      auto fetch = [&](int ch) {
        oh = ow = mraw = 0.f;
        if (ch < c_end) {
          const int b = ch / per_img, rr = ch % per_img;
          const int ho = (rr / p.tiles_w) * p.TH + (r >> p.tw_shift), wo = (rr % p.tiles_w) * p.TW + (r & (p.TW - 1));
          if (ho < g.Ho && wo < g.Wo) {
            const long long pl = (static_cast<long long>(b) * g.Ho + ho) * g.Wo + wo;
            oh = __ldg(p.offset + pl * g.ldo + 2 * tap);
            ow = __ldg(p.offset + pl * g.ldo + 2 * tap + 1);
            if (p.mask) mraw = __ldg(p.mask + pl * g.ldm + tap);
          }
        }
      };
      uint32_t npi[4];
      float nwt[4];
      auto geom = [&](int ch) {          // from the fetched (oh, ow, mraw) of patch ch
#pragma unroll
        for (int q = 0; q < 4; ++q) { npi[q] = 0u; nwt[q] = 0.f; }
        if (ch < c_end) {
          const int b = ch / per_img, rr = ch % per_img;
          const int ho = (rr / p.tiles_w) * p.TH + (r >> p.tw_shift), wo = (rr % p.tiles_w) * p.TW + (r & (p.TW - 1));
          if (ho < g.Ho && wo < g.Wo) {
            // mul then add, each rounded (sample_pos)
            const float h = __fadd_rn(__fmul_rn(static_cast<float>(ho * g.sh - g.ph + ti * g.dh), g.scale_h), oh);
            const float w = __fadd_rn(__fmul_rn(static_cast<float>(wo * g.sw - g.pw + tj * g.dw), g.scale_w), ow);
            corner_pixels(g, b, h, w, npi, nwt);
            float m = 1.f;
            if (p.mask) m = g.mask_logits ? 1.f / (1.f + __expf(-mraw)) : mraw;
#pragma unroll
            for (int q = 0; q < 4; ++q) nwt[q] *= m;
          }
        }
      };
      uint4 ld[2][4];
#define LSN_WISSUE(buf, cb_, pi_)                                                                   \
  {                                                                                                 \
    const unsigned long long xc_ = xt + (cb_) * (FK * 2);                                           \
    ld[buf][0] = ldg128_at(xc_, pi_[0], ldxb);                                                      \
    ld[buf][1] = ldg128_at(xc_, pi_[1], ldxb);                                                      \
    ld[buf][2] = ldg128_at(xc_, pi_[2], ldxb);                                                      \
    ld[buf][3] = ldg128_at(xc_, pi_[3], ldxb);                                                      \
  }
#define LSN_WCONSUME(buf, cb_)                                                                      \
  {                                                                                                 \
    uint32_t outw_[4];                                                                              \
    _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) {                                              \
      const uint32_t a0 = reinterpret_cast<const uint32_t*>(&ld[buf][0])[i_];                       \
      const uint32_t a1 = reinterpret_cast<const uint32_t*>(&ld[buf][1])[i_];                       \
      const uint32_t a2 = reinterpret_cast<const uint32_t*>(&ld[buf][2])[i_];                       \
      const uint32_t a3 = reinterpret_cast<const uint32_t*>(&ld[buf][3])[i_];                       \
      float2 acc_ = __fmul2_rn(make_float2(cw[0], cw[0]), bf16x2_f2(a0));                           \
      acc_ = __ffma2_rn(make_float2(cw[1], cw[1]), bf16x2_f2(a1), acc_);                            \
      acc_ = __ffma2_rn(make_float2(cw[2], cw[2]), bf16x2_f2(a2), acc_);                            \
      acc_ = __ffma2_rn(make_float2(cw[3], cw[3]), bf16x2_f2(a3), acc_);                            \
      outw_[i_] = pack_bf16x2(acc_.x, acc_.y);                                                      \
    }                                                                                               \
    sts128(sB + (cb_) * (WK * 128), make_uint4(outw_[0], outw_[1], outw_[2], outw_[3]));            \
  }
      fetch(c_begin);
      geom(c_begin);
      fetch(c_begin + 1);
      uint32_t cpi[4];
      float cw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { cpi[q] = npi[q]; cw[q] = nwt[q]; }
      LSN_WISSUE(0, 0, cpi)
      LSN_WISSUE(1, 1, cpi)
      geom(c_begin + 1);
      fetch(c_begin + 2);
      for (int ch = c_begin; ch < c_end; ++ch) {
        const bool more = ch + 1 < c_end;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t sB = smem_s + stage * W_STAGE + W_ABYTES + row_off;
        LSN_WCONSUME(0, 0)
        LSN_WISSUE(0, 2, cpi)
        LSN_WCONSUME(1, 1)
        LSN_WISSUE(1, 3, cpi)
        LSN_WCONSUME(0, 2)
        if (more) LSN_WISSUE(0, 0, npi)
        LSN_WCONSUME(1, 3)
        if (more) LSN_WISSUE(1, 1, npi)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
        for (int q = 0; q < 4; ++q) { cpi[q] = npi[q]; cw[q] = nwt[q]; }
        geom(ch + 2);
        fetch(ch + 3);
      }
#undef LSN_WISSUE
#undef LSN_WCONSUME
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[i] += sum_s part[s * stride + i] in split order (deterministic second stage of the weight gradient)
__global__ void dcn_wgrad_reduce_kernel(const float* __restrict__ part, long long stride, int splits, float* __restrict__ out,
                                        long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 acc = reinterpret_cast<const float4*>(out)[i];
  for (int s = 0; s < splits; ++s) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + s * stride) + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// 0: never, 1: whenever the shape is supported, 2 (default): where it measured faster than gather -> GEMM.
// Env LSNET_DCN_FUSED; lsnet_dcn_fused_enable() overrides it at run time.
static int g_fused_on = -1;
void dcn_fused_set(int on) { g_fused_on = on < 0 ? 2 : (on > 2 ? 2 : on); }

// Shapes the fused kernels take; everything else goes through the column-matrix path.
bool dcn_fused_supported(int C, int N, int kh, int kw, int dg, long long ldx, long long B, long long H, long long W,
                         long long Ho, long long Wo) {
  if (g_fused_on < 0) dcn_fused_set(env_int("LSNET_DCN_FUSED", 2));
  if (!g_fused_on) return false;
  const bool ok = dg == 1 && kh * kw <= F_MAXTAPS && C % FK == 0 && N >= 16 && N % 16 == 0 && N <= 256 &&
                  ldx % 8 == 0 && ldx * 2 < 2147483647LL && B * H * W < 2147483647LL && B * Ho * Wo < 2147483647LL;
  if (!ok || g_fused_on == 1) return ok;
  // auto: a 128-pixel patch runs its 9 * C/64 K blocks back to back on one SM (~1 us each), so the fused kernel only
  // wins once the patches fill the machine at least twice (r02, B200: 100x168x4 -> 0.145 vs 0.181 ms; 50x84x4 -> 0.080
  // vs 0.060 ms; whole step 23.9 ms with every level fused vs 23.6 ms without)
  int TH = 8, TW = 16;
  pick_patch(static_cast<int>(Ho), static_cast<int>(Wo), FM, &TH, &TW);
  const long long tiles = B * ((Ho + TH - 1) / TH) * ((Wo + TW - 1) / TW);
  return tiles >= 2LL * num_sms();
}

template <int BN, int STAGES, bool SAVE_COL>
static int launch_fused_fwd(const CUtensorMap& tmB, const FusedFwdArgs& a, cudaStream_t st) {
  using Cfg = FCfg<BN, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dcn_fused_fwd_kernel<BN, STAGES, SAVE_COL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(dcn_fused_fwd<%d,%d>): %s", BN, STAGES, cudaGetErrorString(e));
    attr_done = true;
  }
  const long long items = static_cast<long long>(a.num_tiles) * a.cb_parts;
  const int grid = items < num_sms() ? static_cast<int>(items) : num_sms();
  const double px = static_cast<double>(a.g.B) * a.g.Ho * a.g.Wo;
  const int th = timing_begin(TC_DCN_FWD, 2.0 * px * a.N * a.taps * (a.grouped ? FK : a.g.C), st);
  dcn_fused_fwd_kernel<BN, STAGES, SAVE_COL><<<grid, F_THREADS, Cfg::kSmemBytes, st>>>(tmB, a);
  timing_end(th, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("dcn_fused_fwd<%d,%d> launch: %s", BN, STAGES, cudaGetErrorString(e));
  count_launch();
  return 0;
}

template <int BN>
static int dispatch_fused_fwd(const CUtensorMap& tmB, const FusedFwdArgs& a, cudaStream_t st) {
  static int stages = 0;
  if (!stages) {
    stages = env_int("LSNET_DCN_FUSED_STAGES", 3);      // in-step: 3 stages 23.9 ms, 2 stages 24.2 ms
    if (stages != 2 && stages != 3) stages = 3;
  }
  if (a.col) return stages == 3 ? launch_fused_fwd<BN, 3, true>(tmB, a, st) : launch_fused_fwd<BN, 2, true>(tmB, a, st);
  return stages == 3 ? launch_fused_fwd<BN, 3, false>(tmB, a, st) : launch_fused_fwd<BN, 2, false>(tmB, a, st);
}

// grouped != 0: Wp is the block-diagonal pack [N = C, taps * 64] (see lsnet_dcn_forward)
int dcn_fused_forward(const DcnGeom& g, const void* x, const float* offset, const float* mask, const void* Wp, int N,
                      const float* bias, int relu, void* out, long long ldc, int out_fp32, void* col, cudaStream_t st,
                      int grouped, double* gn_sums, int gn_G) {
  FusedFwdArgs a{};
  a.gn_sums = gn_sums; a.gn_G = gn_G; a.gn_cpg = gn_sums ? N / gn_G : 0;
  a.grouped = grouped ? 1 : 0;
  a.g = g;
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.offset = offset;
  a.mask = mask;
  a.N = N;
  a.taps = g.kh * g.kw;
  a.cblks = g.C / FK;
  int TH = 8, TW = 16;
  pick_patch(g.Ho, g.Wo, FM, &TH, &TW);
  a.TH = TH; a.TW = TW;
  a.tw_shift = 0;
  while ((1 << a.tw_shift) < TW) ++a.tw_shift;
  a.tiles_h = (g.Ho + TH - 1) / TH;
  a.tiles_w = (g.Wo + TW - 1) / TW;
  a.num_tiles = g.B * a.tiles_h * a.tiles_w;
  a.cb_parts = 1;
  a.cb_per_part = a.cblks;
  if (a.grouped) {      // deal the independent channel blocks of a patch to several work items until the SMs are covered twice
    while (a.cb_per_part > 1 && static_cast<long long>(a.num_tiles) * a.cb_parts < 2LL * num_sms()) {
      a.cb_per_part = (a.cb_per_part + 1) / 2;
      a.cb_parts = (a.cblks + a.cb_per_part - 1) / a.cb_per_part;
    }
  }
  a.out = out; a.ldc = ldc; a.out_fp32 = out_fp32; a.relu = relu; a.bias = bias;
  a.col = static_cast<__nv_bfloat16*>(col);
  const int BN = grouped ? 64 : (N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32)));
  const long long K = grouped ? static_cast<long long>(a.taps) * FK : static_cast<long long>(a.taps) * g.C;
  CUtensorMap tmB;
  if (int rc = make_map_2d(&tmB, Wp, N, K, K, 64, BN)) return rc;
  switch (BN) {
    case 256: return dispatch_fused_fwd<256>(tmB, a, st);
    case 128: return dispatch_fused_fwd<128>(tmB, a, st);
    case 64: return dispatch_fused_fwd<64>(tmB, a, st);
    default: return dispatch_fused_fwd<32>(tmB, a, st);
  }
}


// Weight gradient with re-sampled columns.  `partial` (optional, fp32 [splits, M, ldw] from
// dcn_fused_wgrad_partial_bytes) selects the deterministic two-stage reduction.
static int wgrad_splits(const DcnGeom& g, int M, int* k_chunks, int* TH, int* TW, int diag = 0) {
  pick_patch(g.Ho, g.Wo, WK, TH, TW);
  *k_chunks = g.B * ((g.Ho + *TH - 1) / *TH) * ((g.Wo + *TW - 1) / *TW);
  const int base = g.kh * g.kw * ((g.C + 255) / 256) * (diag ? 1 : (M + 255) / 256);
  int splits = (num_sms() + base / 2) / base;          // ~ one item per SM
  if (splits > *k_chunks) splits = *k_chunks;
  if (splits < 1) splits = 1;
  return splits;
}

size_t dcn_fused_wgrad_partial_bytes(const DcnGeom& g, int M, long long ldw, int diag) {
  int kc, TH, TW;
  const int splits = wgrad_splits(g, M, &kc, &TH, &TW, diag);
  return static_cast<size_t>(splits) * M * ldw * sizeof(float);
}

// diag != 0 (grouped weights, M == C): dW is the compact [M, taps * 256] block-diagonal form (ldw = taps * 256)
int dcn_fused_wgrad(const DcnGeom& g, const void* dy, long long ldy, int M, const void* x, const float* offset,
                    const float* mask, float* dW, long long ldw, float* partial, cudaStream_t st, int diag) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(dcn_fused_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(dcn_fused_wgrad): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  FusedWgradArgs a{};
  a.g = g;
  a.x = static_cast<const __nv_bfloat16*>(x);
  a.offset = offset;
  a.mask = mask;
  a.M = M;
  int TH = 4, TW = 16, kc = 0;
  a.diag = diag ? 1 : 0;
  a.splits = wgrad_splits(g, M, &kc, &TH, &TW, diag);
  a.k_chunks = kc;
  a.TH = TH; a.TW = TW;
  a.tw_shift = 0;
  while ((1 << a.tw_shift) < TW) ++a.tw_shift;
  a.tiles_h = (g.Ho + TH - 1) / TH;
  a.tiles_w = (g.Wo + TW - 1) / TW;
  a.taps = g.kh * g.kw;
  a.n_tiles = (g.C + 255) / 256;
  a.m_groups = diag ? 1 : (M + 255) / 256;
  a.chunks_per_split = (kc + a.splits - 1) / a.splits;
  a.ldw = ldw;
  if (partial) {
    a.out = partial;
    a.split_stride = static_cast<long long>(M) * ldw;
    // splits whose pixel range is empty never store: the reduce must not read garbage
    cudaMemsetAsync(partial, 0, static_cast<size_t>(a.splits) * M * ldw * sizeof(float), st);
  } else {
    a.out = dW;
    a.split_stride = 0;
  }
  CUtensorMap tmA;
  if (int rc = make_map_nhwc(&tmA, dy, g.B, g.Ho, g.Wo, M, ldy, TW, TH)) return rc;
  const int items = a.taps * a.n_tiles * a.m_groups * a.splits;
  const int grid = items < num_sms() ? items : num_sms();
  const double px = static_cast<double>(g.B) * g.Ho * g.Wo;
  const int th = timing_begin(TC_DCN_WGRAD, 2.0 * px * M * a.taps * (diag ? 256 : g.C), st);
  dcn_fused_wgrad_kernel<<<grid, F_THREADS, W_SMEM, st>>>(tmA, a);
  if (partial) {
    const long long n4 = static_cast<long long>(M) * ldw / 4;
    dcn_wgrad_reduce_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(partial, a.split_stride, a.splits, dW, n4);
    count_launch();
  }
  timing_end(th, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("dcn_fused_wgrad launch: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace lsn
