// tcgen05 / TMEM / TMA GEMM family for the LSNet hot path (sm_100a only).
//
//  gemm_kmajor<BN>  C[M,N] = A[M,K] . B[N,K]^T (+bias, ReLU), bf16 in, fp32 accumulate in TMEM.
//     A is either a plain row-major [M,K] matrix (DCN column matrix, 1x1 convs, dY for bwd-data) or -- conv mode --
//     an NHWC activation read through a 4-D tensor map with one shifted TMA box per filter tap (implicit GEMM,
//     zero padding = TMA out-of-bounds fill).  Replaces reference K2/K3 (deform_conv_cuda.cpp:673-691) and the
//     cuDNN convs of FPN / LSHead (necks/fpn.py:165-217, dense_heads/lsnet_head.py:502-755).
//  gemm_mnmajor     C[M,N] += A[K,M]^T . B[K,N]  (reduction over rows = pixels), both operands MN-major in smem,
//     split-K over pixel ranges with fp32 red.global.add.  Replaces the weight-gradient GEMMs
//     (deform_conv_cuda.cpp:782-787, 1113-1124) and conv weight grads.
//
// Warp roles (320 threads): warp0 = TMA producer, warp1 = TMEM allocator + single-thread MMA issuer,
// warps2-9 = epilogue (warp w owns TMEM lane quarter w%4; the two warps of a quarter split the tile's columns).  Persistent over output tiles; smem ring of
// kStages {A,B} tiles; two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <stdlib.h>

#include "common.cuh"
#include "lsnet_internal.h"

namespace lsn {

constexpr int BM = 128;       // UMMA M (rows of the accumulator = TMEM lanes)
constexpr int BK = 64;        // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kThreads = 320;   // warp0 TMA, warp1 MMA, warps2-9 epilogue (two warps per TMEM lane quarter)

constexpr int kMaxTaps = 16;

struct GemmArgs {
  int M, N;            // logical output extent (N multiple of 16)
  int num_k_iters;     // K/64, or taps * C/64 in conv mode
  int m_tiles, n_tiles;
  // conv mode (A through a 4-D NHWC map); ignored when conv == 0.  The GEMM rows are the pixels of a logical Ho x Wo grid
  // (tiled TH x TW); filter tap t reads input pixel (h*ish + tap_dy[t], w*isw + tap_dx[t]) -- the map's element strides
  // are (ish, isw), out-of-range coordinates are TMA zero fill -- and the row is written to pixel
  // (h*osh + ooh, w*osw + oow) of an OH x OW output grid (osh = 1, ooh = 0: plain; osh = 2: one phase of the input
  // gradient of a stride-2 convolution).
  int conv, Ho, Wo, tiles_h, tiles_w, TH, TW, tw_shift, cblks, ish, isw, osh, osw, ooh, oow, OH, OW;   // TW = 1 << tw_shift
  signed char tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_kb[kMaxTaps];   // tap_kb: K block (of cblks*64 columns) of the weight pack
  void* out;           // [M, ldc] bf16 or fp32
  long long ldc;
  int out_fp32, relu;
  const float* bias;   // [N] or null
  // epilogue extras (same row mapping as out):  v = acc + bias + resid;  relu;  v = mask > 0 ? v : 0
  const __nv_bfloat16* resid;
  long long ldr;
  const __nv_bfloat16* mask;
  long long ldm;
  // GroupNorm statistics of the output (before ReLU), accumulated by the epilogue: sums [B, G, 2] fp64 (pre-zeroed), gn_cpg
  // channels per group, gn_hw pixels per image (row -> image in plain mode); null = off
  double* gn_sums;
  int gn_cpg, gn_G;
  long long gn_hw;
  // block-diagonal mode (grouped DCN dCol): ONE K block per tile; the N tile nt (64 columns) multiplies the 64-column
  // block (nt % blockdiag) of A with B[n0 .. n0+63, 0..63]; 0 = off
  int blockdiag;
};

// OCC = CTAs per SM.  OCC = 2 (BN <= 128: 2 x 2*BN TMEM columns, <= 113 KB of shared memory each, <= 102 registers per thread)
// is for launches below the FLOP/byte ridge: their tiles have 1-4 K steps, so a CTA is mostly its epilogue -- a serial
// latency chain per warp (ncu: 16 % warps active, 28 % issue slots with one CTA of 10 warps per SM) -- and a second
// resident CTA hides it.
template <int BN, int OCC = 1>
struct KCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = OCC == 1 ? ((BN == 256) ? 4 : (BN == 128 ? 6 : 8)) : (BN == 128 ? 2 : (BN == 64 ? 3 : 4));
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kStagingBytes = 8 * 32 * 80;   // per epilogue warp: 32 rows x (64 B + 16 B pad)
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kEpiWarps = BN >= 64 ? 8 : 4;        // column halves only pay off from 64 columns up
  static constexpr int kColsPerWarp = BN >= 64 ? BN / 2 : BN;
  static_assert(OCC == 1 || (kSmemBytes <= 113 * 1024 && OCC * kTmemCols <= 512), "two CTAs per SM must fit");
};

template <int BN, int OCC>
__global__ void __launch_bounds__(kThreads, OCC)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmArgs p) {
  using Cfg = KCfg<BN, OCC>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint32_t* staging = reinterpret_cast<uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], Cfg::kEpiWarps);   // one elected lane per epilogue warp
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

  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
        const int n0 = nt * BN;
        int m0 = mt * BM, b = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int per_img = p.tiles_h * p.tiles_w;
          b = mt / per_img;
          const int r = mt % per_img;
          h0 = (r / p.tiles_w) * p.TH;
          w0 = (r % p.tiles_w) * p.TW;
        }
        for (int it = 0; it < p.num_k_iters; ++it) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          int kb = it;
          if (p.conv) {
            const int tap = it / p.cblks, cb = it % p.cblks;
            kb = p.tap_kb[tap] * p.cblks + cb;
            tma_load_4d(sA, &tmA, &full_bar[stage], cb * BK, w0 * p.isw + p.tap_dx[tap], h0 * p.ish + p.tap_dy[tap], b);
          } else {
            tma_load_2d(sA, &tmA, &full_bar[stage], p.blockdiag ? (nt % p.blockdiag) * BK : it * BK, m0);
          }
          tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int it = 0; it < p.num_k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sB = sA + Cfg::kABytes;
          const uint64_t adesc = umma_desc_sw128(sA, 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(sB, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B along K inside the 128-B swizzle row = +2 in the (addr>>4) field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                    // frees the smem slot when these MMAs retire
          if (it == p.num_k_iters - 1) umma_commit(&tfull_bar[as]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp - 2 < Cfg::kEpiWarps) {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;     // accumulator row handled by this thread
    const int col_begin = ((warp - 2) >> 2) * Cfg::kColsPerWarp;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / p.n_tiles, nt = tile % p.n_tiles;
      const int n0 = nt * BN;
      // Output row (and validity) of accumulator row rr of this tile.  The tile's origin is decoded ONCE per tile (the
      // per-row part is shifts / masks: TW is a power of two); the rows this thread stores in the transposed write-out
      // are precomputed too -- with the divisions inside the column loop the epilogue of a conv tile was ALU bound.
      int tb = 0, th0 = 0, tw0 = 0;
      if (p.conv) {
        const int per_img = p.tiles_h * p.tiles_w;
        tb = mt / per_img;
        const int t2 = mt - tb * per_img;
        const int ty = t2 / p.tiles_w;
        th0 = ty * p.TH;
        tw0 = (t2 - ty * p.tiles_w) * p.TW;
      }
      auto row_of = [&](int rr, bool* ok) -> long long {
        if (p.conv) {
          const int h = th0 + (rr >> p.tw_shift), w = tw0 + (rr & (p.TW - 1));
          const int oh = h * p.osh + p.ooh, ow = w * p.osw + p.oow;
          *ok = (h < p.Ho) && (w < p.Wo) && (oh < p.OH) && (ow < p.OW);
          return (static_cast<long long>(tb) * p.OH + oh) * p.OW + ow;
        }
        const long long g = static_cast<long long>(mt) * BM + rr;
        *ok = g < p.M;
        return g;
      };
      bool valid;
      const long long row = row_of(r, &valid);
      long long srow[4];
      bool sok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) srow[i] = row_of(q * 32 + (lane >> 2) + 8 * i, &sok[i]);
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      // 32-bit shared-space address of this warp's staging rows (80-byte pitch): explicit ld/st.shared -- through a generic
      // pointer these compile to LD.E / ST.E whose long-scoreboard latency sat on every store of the epilogue (ncu)
      const uint32_t stg_s = smem_u32(staging + (warp - 2) * (32 * 20));
#pragma unroll 1
      for (int c = col_begin; c < col_begin + Cfg::kColsPerWarp; c += 32) {
        const int col0 = n0 + c;
        if (col0 >= p.N) continue;     // warp-uniform (this warp neither reads the accumulator columns nor stores them)
        // every global operand of the epilogue is requested BEFORE the accumulator is awaited, as independent loads: issued
        // one by one next to their uses they cost eight serial L1/L2 round trips per 32-column block (ncu: 40 % of the
        // warp-stall samples of a K = 64 layer sat on the bias adds)
        float4 bv[8];
        uint4 rv[4], mv[4];
        if (p.bias) {      // N % 16 == 0 and the bias vector is 16-byte aligned (host-checked): whole float4 groups
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bv[j] = (col0 + 4 * j < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // resid / mask: requested in the COALESCED pattern of the write-out (8 rows x 64 contiguous bytes per warp
        // instruction) and transposed to row-per-thread through the warp's staging buffer below -- one thread reading
        // its own 64-byte row piece touches 32 different lines per instruction and doubled the kernel's time
        const int seg_l = lane & 3;
        if (p.resid) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            rv[i] = (sok[i] && col0 + seg_l * 8 < p.N)
                        ? __ldg(reinterpret_cast<const uint4*>(p.resid + srow[i] * p.ldr + col0 + seg_l * 8))
                        : make_uint4(0u, 0u, 0u, 0u);
        }
        if (p.mask) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            mv[i] = (sok[i] && col0 + seg_l * 8 < p.N)
                        ? __ldg(reinterpret_cast<const uint4*>(p.mask + srow[i] * p.ldm + col0 + seg_l * 8))
                        : make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
        }
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[4 * j] += bv[j].x; f[4 * j + 1] += bv[j].y; f[4 * j + 2] += bv[j].z; f[4 * j + 3] += bv[j].w;
          }
        }
        if (p.resid) {
          // transpose: (row (lane>>2)+8i, 16-byte segment lane&3) -> this thread's row, segments 0..3
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(stg_s + (((lane >> 2) + 8 * i) * 20 + seg_l * 4) * 4, rv[i]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) rv[j] = lds128(stg_s + (lane * 20 + j * 4) * 4);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t rw[4] = {rv[j].x, rv[j].y, rv[j].z, rv[j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              f[8 * j + 2 * i] += __uint_as_float(rw[i] << 16);
              f[8 * j + 2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
            }
          }
        }
        if (p.gn_sums) {
          int b = tb, b_lo = tb, b_hi = tb;
          if (!p.conv) {      // plain [pixels, C] rows: the 32 rows of this warp may straddle an image boundary
            const long long r0 = static_cast<long long>(mt) * BM + q * 32;
            const long long rl = (r0 + 31 < p.M ? r0 + 31 : p.M - 1);
            b = static_cast<int>(row / p.gn_hw);
            b_lo = static_cast<int>(r0 / p.gn_hw);
            b_hi = static_cast<int>(rl / p.gn_hw);
          }
          gn_epilogue_sums(f, valid, b, b_lo, b_hi, col0, p.N, p.gn_cpg, p.gn_G, p.gn_sums, lane);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.mask) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(stg_s + (((lane >> 2) + 8 * i) * 20 + seg_l * 4) * 4, mv[i]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) mv[j] = lds128(stg_s + (lane * 20 + j * 4) * 4);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t mw[4] = {mv[j].x, mv[j].y, mv[j].z, mv[j].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {      // bf16 > 0  <=>  sign clear and not (+)zero
              if (!((mw[i] & 0xffffu) - 1u < 0x7fffu)) f[8 * j + 2 * i] = 0.f;
              if (!((mw[i] >> 16) - 1u < 0x7fffu)) f[8 * j + 2 * i + 1] = 0.f;
            }
          }
        }
        if (p.out_fp32) {
          if (valid) {
            float* o = reinterpret_cast<float*>(p.out) + row * p.ldc + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (col0 + j < p.N)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
        } else {
          // stage the 32x32 bf16 block in shared memory (row pitch 80 B: conflict-free 16-byte accesses), then store
          // it with 8 rows x 64 contiguous bytes per warp instruction instead of 32 rows x 16 bytes
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
            if (sok[i] && col0 + seg * 8 < p.N) {
              const uint4 val = lds128(stg_s + (rr * 20 + seg * 4) * 4);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + srow[i] * p.ldc + col0 + seg * 8) = val;
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

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------
// MN-major split-K GEMM:  C[M,N] += sum_{p in split} A[p, m] * B[p, n]   (A: [P, M] row-major, B: [P, N] row-major)
// In conv mode B rows are the NHWC pixels shifted by the filter tap (4-D map), A rows the dY pixels (4-D map, no
// shift); the K "pixel" loop walks TH x TW = 64-pixel patches.  Output fp32 with red.global.add (C pre-zeroed).
// ---------------------------------------------------------------------------------------------------------
struct WgradArgs {
  int M, N;              // M = C_out, N = columns of B handled per tap (C_in or 9*C for DCN columns)
  int m_tiles, n_tiles;  // tiles of 128 x BNW
  int taps;              // 1 (plain) or kh*kw (conv)
  int splits;            // split-K factor over pixel chunks
  int k_chunks;          // total 64-pixel chunks (plain: ceil(P/64); conv: B * tiles_h * tiles_w)
  int conv, tiles_h, tiles_w, TH, TW, kw, pad_h, pad_w, dil_h, dil_w, ish, isw;   // (ish, isw): stride of the convolution
  float* out;            // [M, taps, N] fp32 (row stride ldc = taps*N)
  long long ldc;
};

constexpr int BNW = 256;

// MT = number of 128-row output tiles (of C_out) one CTA accumulates at once.  MT = 2 keeps a 256 x 256 fp32
// accumulator in all 512 TMEM columns and loads the B (pixel x C_in) tile ONCE for both halves: the kernel is
// L2-bandwidth bound (every 64-pixel chunk is re-read by each output tile that needs it), so doubling the tile
// height cuts the L2 traffic per FLOP by a third.  MT = 1 double-buffers the accumulator instead.
template <int MT>
struct WCfg {
  static constexpr int kABytes = BK * BM * 2 * MT;    // 64 pixels x (MT*128) couts
  static constexpr int kBBytes = BK * BNW * 2;        // 64 pixels x 256 cins
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = MT == 1 ? 4 : 3;
  static constexpr int kAccStages = MT == 1 ? 2 : 1;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int MT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_mnmajor_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const WgradArgs p) {
  using Cfg = WCfg<MT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);
    }
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

  // work item = (tap, m_group, n_tile, split); m_group covers MT consecutive 128-row tiles
  const int m_groups = (p.m_tiles + MT - 1) / MT;
  const int items = p.taps * m_groups * p.n_tiles * p.splits;
  const int chunks_per_split = (p.k_chunks + p.splits - 1) / p.splits;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int t = item;
        const int split = t % p.splits; t /= p.splits;
        const int nt = t % p.n_tiles; t /= p.n_tiles;
        const int mg = t % m_groups; t /= m_groups;
        const int tap = t;
        const int c_begin = split * chunks_per_split;
        const int c_end = min(p.k_chunks, c_begin + chunks_per_split);
        const int dy = p.conv ? tap / p.kw : 0, dx = p.conv ? tap % p.kw : 0;
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (p.conv) {
            const int per_img = p.tiles_h * p.tiles_w;
            const int b = ch / per_img, rr = ch % per_img;
            const int h0 = (rr / p.tiles_w) * p.TH, w0 = (rr % p.tiles_w) * p.TW;
#pragma unroll
            for (int j = 0; j < MT * BM / 64; ++j)
              tma_load_4d(sA + j * (BK * 128), &tmA, &full_bar[stage], mg * MT * BM + j * 64, w0, h0, b);
#pragma unroll
            for (int j = 0; j < BNW / 64; ++j)
              tma_load_4d(sB + j * (BK * 128), &tmB, &full_bar[stage], nt * BNW + j * 64,
                          w0 * p.isw - p.pad_w + dx * p.dil_w, h0 * p.ish - p.pad_h + dy * p.dil_h, b);
          } else {
#pragma unroll
            for (int j = 0; j < MT * BM / 64; ++j)
              tma_load_2d(sA + j * (BK * 128), &tmA, &full_bar[stage], mg * MT * BM + j * 64, ch * BK);
#pragma unroll
            for (int j = 0; j < BNW / 64; ++j)
              tma_load_2d(sB + j * (BK * 128), &tmB, &full_bar[stage], nt * BNW + j * 64, ch * BK);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BNW, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int split = item % p.splits;
        const int c_begin = split * chunks_per_split;
        const int c_end = min(p.k_chunks, c_begin + chunks_per_split);
        if (c_end <= c_begin) continue;   // empty split: epilogue skips it too
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BNW);
        for (int ch = c_begin; ch < c_end; ++ch) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sB = sA + Cfg::kABytes;
          // MN-major: LBO = distance between 64-wide M/N chunks (one TMA box = 64 rows x 128 B), SBO = 8 K rows
          const uint64_t bdesc = umma_desc_sw128(sB, BK * 128, 1024);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = umma_desc_sw128(sA + mt * 2 * (BK * 128), BK * 128, 1024);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // 16 K rows = 2 swizzle atoms of 8 rows = +2048 B -> +128 in the (addr>>4) field
              umma_bf16(tmem_d + mt * BNW, adesc + 128 * k, bdesc + 128 * k, idesc, ((ch - c_begin) | k) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (ch == c_end - 1) umma_commit(&tfull_bar[as]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == Cfg::kAccStages) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int col_begin = ((warp - 2) >> 2) * (BNW / 2);
    int as = 0;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int t = item;
      const int split = t % p.splits; t /= p.splits;
      const int nt = t % p.n_tiles; t /= p.n_tiles;
      const int mg = t % m_groups; t /= m_groups;
      const int tap = t;
      const int c_begin = split * chunks_per_split;
      const int c_end = min(p.k_chunks, c_begin + chunks_per_split);
      if (c_end <= c_begin) continue;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int m = (mg * MT + mt) * BM + r;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(as * BNW + mt * BNW);
#pragma unroll 1
        for (int c = col_begin; c < col_begin + BNW / 2; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
          const int col0 = nt * BNW + c;
          if (m < p.M && col0 < p.N) {
            float* o = p.out + static_cast<long long>(m) * p.ldc + static_cast<long long>(tap) * p.N + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < p.N) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j),
                             "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                             "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                             : "memory");
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == Cfg::kAccStages) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =========================================================================================================
// Host side: tensor maps + launch
// =========================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  // cuTensorMapEncodeTiled needs a current context: autograd worker threads may reach here before any runtime call bound
  // the primary context to them (CUDA_ERROR_INVALID_CONTEXT otherwise)
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// rank-2 bf16 map over a row-major [rows, cols] matrix with row pitch ld (elements); box = [box_cols, box_rows]
int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                       uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu ld=%llu)", (int)r,
                                          (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return 0;
}
// rank-4 bf16 map over NHWC [B,H,W,C] (pixel pitch ldp elements); box = [64, TW, TH, 1] pixels taken every (sh, sw)-th
// row / column (TMA element strides: the box spans TW*sw x TH*sh input pixels and lands compacted in shared memory)
int make_map_nhwc(CUtensorMap* m, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint64_t ldp,
                  uint32_t TW, uint32_t TH, uint32_t sw, uint32_t sh) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  if (TW * sw > 256 || TH * sh > 256) return set_error("make_map_nhwc: box %ux%u with strides %ux%u exceeds 256", TW, TH, sw, sh);
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {ldp * 2, W * ldp * 2, H * W * ldp * 2};
  cuuint32_t box[4] = {64, TW * sw, TH * sh, 1};
  cuuint32_t estr[4] = {1, sw, sh, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(4d) failed: %d", (int)r);
  return 0;
}

// rank-5 bf16 SWIZZLE_128B map over a DCN column / dCol matrix [B*Ho*Wo, taps*C] viewed as [B, Ho, Wo, taps, C] (row pitch
// ldcol elements); box = [64 channels, taps, PW, PH, 1]: one TMA instruction stages the 64-channel slice of a pixel patch
int make_map_col5d(CUtensorMap* m, const void* ptr, uint64_t B, uint64_t Ho, uint64_t Wo, uint64_t taps, uint64_t C,
                   uint64_t ldcol, uint32_t PW, uint32_t PH) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[5] = {C, taps, Wo, Ho, B};
  cuuint64_t strides[4] = {C * 2, ldcol * 2, Wo * ldcol * 2, Ho * Wo * ldcol * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(taps), PW, PH, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("cuTensorMapEncodeTiled(5d) failed: %d", (int)r);
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// choose a spatial patch TH x TW = pixels with the least padding waste
void pick_patch(int H, int W, int pixels, int* TH, int* TW) {
  long long best = -1;
  for (int th = 1; th <= pixels; th <<= 1) {
    int tw = pixels / th;
    if (tw > 256 || th > 256) continue;
    long long cover = static_cast<long long>((H + th - 1) / th) * th * ((W + tw - 1) / tw) * tw;
    if (best < 0 || cover < best) { best = cover; *TH = th; *TW = tw; }
  }
}

// FLOPs / algorithmic bytes (input read once, weights, output, epilogue operands) of a launch against the ridge of the
// machine (measured 1346.8 TFLOP/s / 6.54 TB/s ~ 200 FLOP/B)
static bool gemm_hbm_bound(const GemmArgs& a, double* flops_out, double* bytes_out) {
  const double flops = 2.0 * a.M * a.N * a.num_k_iters * BK;
  const double k_in = a.conv ? static_cast<double>(a.cblks) * BK * ((a.num_k_iters > a.cblks) ? a.ish * a.isw : 1)
                             : static_cast<double>(a.num_k_iters) * BK;
  const double bytes = 2.0 * a.M * k_in + 2.0 * a.N * a.num_k_iters * BK + static_cast<double>(a.M) * a.N * (a.out_fp32 ? 4 : 2) +
                       (a.resid ? 2.0 * a.M * a.N : 0.0) + (a.mask ? 2.0 * a.M * a.N : 0.0);
  if (flops_out) *flops_out = flops;
  if (bytes_out) *bytes_out = bytes;
  return flops < 200.0 * bytes;
}

template <int BN, int OCC>
static int launch_kmajor(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
  using Cfg = KCfg<BN, OCC>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kmajor_kernel<BN, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(gemm_kmajor<%d,%d>): %s", BN, OCC, cudaGetErrorString(e));
    attr_done = true;
  }
  int tiles = a.m_tiles * a.n_tiles;
  int grid = tiles < OCC * num_sms() ? tiles : OCC * num_sms();
  double flops, bytes;
  const bool hbm_bound = gemm_hbm_bound(a, &flops, &bytes);
  const int th = timing_begin(hbm_bound ? TC_GEMM_HBM : TC_GEMM, hbm_bound ? bytes : flops, st);
  gemm_kmajor_kernel<BN, OCC><<<grid, kThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, a);
  timing_end(th, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm_kmajor<%d,%d> launch: %s", BN, OCC, cudaGetErrorString(e));
  count_launch();
  return 0;
}

static int dispatch_kmajor(const CUtensorMap& tmA, const void* Bw, int N, int K, long long ldb, GemmArgs& a,
                           cudaStream_t st) {
  // N tile: the widest that fits, narrowed when that saves whole rounds of the persistent grid.  A K step of a
  // 128 x BN tile costs max(BN/2 tensor cycles, (128+BN)/4 cycles of shared-memory operand reads) = 128 / 64 / 48 / 40
  // cycles for BN = 256 / 128 / 64 / 32; a launch costs rounds(BN) times that.  Level-0 / level-1 head GEMMs keep 256
  // (ties go to the wider tile: fewer weight re-reads); the trunk's 3x3 convs on 50x84 maps (168 tiles -> two rounds at 57 %)
  // and the small pyramid levels get narrower tiles.  LSNET_GEMM_ADAPT_BN=0 restores the fixed choice.
  const int bn_max = N >= 256 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
  int BN = bn_max;
  static int adapt = -1;
  if (adapt < 0) { const char* e = getenv("LSNET_GEMM_ADAPT_BN"); adapt = e ? atoi(e) : 1; }
  // launches that do not fill one round keep the wide tile: in the training step they run beside other streams' kernels,
  // and more, narrower CTAs only take SMs away from those (measured in-step: 22.9 ms with, 22.5 ms without)
  // (launches below the FLOP/byte ridge keep the wide tile too -- it reads the A operand once; LSNET_GEMM_ADAPT_BN=2 adapts them as well)
  if (adapt && static_cast<long long>(a.m_tiles) * ((N + bn_max - 1) / bn_max) > num_sms() &&
      (adapt == 2 || !gemm_hbm_bound(a, nullptr, nullptr))) {
    const int sms = num_sms();
    long long best = -1;
    for (int bn = bn_max; bn >= 32; bn /= 2) {
      const long long tiles = static_cast<long long>(a.m_tiles) * ((N + bn - 1) / bn);
      const long long cost = ((tiles + sms - 1) / sms) * (bn == 256 ? 128 : bn == 128 ? 64 : bn == 64 ? 48 : 40);
      if (best < 0 || cost < best) { best = cost; BN = bn; }
    }
  }
  // launches below the ridge with enough tiles for two CTAs per SM: 128- (or 64-) column tiles, two CTAs per SM
  static int occ2 = -1;
  if (occ2 < 0) { const char* e = getenv("LSNET_GEMM_OCC2"); occ2 = e ? atoi(e) : 0; }   // measured r02: l1 64->256 0.096 vs 0.063 ms, whole step 23.2 vs 22.35 ms -> opt-in
  int OCC = 1;
  if (occ2 && N >= 64 && gemm_hbm_bound(a, nullptr, nullptr)) {
    const int bn2 = N > 64 ? 128 : 64;
    if (static_cast<long long>(a.m_tiles) * ((N + bn2 - 1) / bn2) >= 2LL * num_sms()) { BN = bn2; OCC = 2; }
  }
  a.n_tiles = (N + BN - 1) / BN;
  CUtensorMap tmB;
  if (int rc = make_map_2d(&tmB, Bw, N, K, ldb, 64, BN)) return rc;
  if (OCC == 2) return BN == 128 ? launch_kmajor<128, 2>(tmA, tmB, a, st) : launch_kmajor<64, 2>(tmA, tmB, a, st);
  switch (BN) {
    case 256: return launch_kmajor<256, 1>(tmA, tmB, a, st);
    case 128: return launch_kmajor<128, 1>(tmA, tmB, a, st);
    case 64: return launch_kmajor<64, 1>(tmA, tmB, a, st);
    default: return launch_kmajor<32, 1>(tmA, tmB, a, st);
  }
}

}  // namespace lsn

using namespace lsn;

extern "C" int lsnet_gemm_bf16(const void* A, long long lda, const void* Bw, long long ldb, void* out, long long ldc,
                               int M, int N, int K, const float* bias, int relu, int out_fp32, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(bias) % 16) return set_error("lsnet_gemm_bf16: bias must be 16-byte aligned");
  if ((N % 16) || (K % 8) || (lda % 8) || (ldb % 8) || (ldc % (out_fp32 ? 4 : 8)))
    return set_error("lsnet_gemm_bf16: need N%%16==0, K%%8==0 and 16-byte aligned pitches (N=%d K=%d lda=%lld ldb=%lld ldc=%lld)",
                     N, K, lda, ldb, ldc);
  CUtensorMap tmA;
  if (int rc = make_map_2d(&tmA, A, M, K, lda, 64, BM)) return rc;
  GemmArgs a{};
  a.M = M; a.N = N; a.num_k_iters = (K + BK - 1) / BK; a.m_tiles = (M + BM - 1) / BM;
  a.conv = 0; a.out = out; a.ldc = ldc; a.out_fp32 = out_fp32; a.relu = relu; a.bias = bias;
  return dispatch_kmajor(tmA, Bw, N, K, ldb, a, static_cast<cudaStream_t>(stream));
}

// internal: lsnet_gemm_bf16 + GroupNorm statistics of the output in the epilogue (dcn_api.cu, column-matrix forward path)
namespace lsn {
int gemm_bf16_gn(const void* A, long long lda, const void* Bw, long long ldb, void* out, long long ldc, int M, int N, int K,
                 const float* bias, int relu, int out_fp32, double* gn_sums, int gn_G, long long gn_hw, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(bias) % 16) return set_error("gemm_bf16_gn: bias must be 16-byte aligned");
  if ((N % 16) || (K % 8) || (lda % 8) || (ldb % 8) || (ldc % (out_fp32 ? 4 : 8)))
    return set_error("gemm_bf16_gn: need N%%16==0, K%%8==0 and 16-byte aligned pitches (N=%d K=%d)", N, K);
  if (gn_sums && (gn_G < 1 || N % gn_G || (N / gn_G) % 8 || gn_hw < 1))
    return set_error("gemm_bf16_gn: GroupNorm epilogue needs N %% G == 0 and (N / G) %% 8 == 0 (N=%d G=%d)", N, gn_G);
  CUtensorMap tmA;
  if (int rc = make_map_2d(&tmA, A, M, K, lda, 64, BM)) return rc;
  GemmArgs a{};
  a.M = M; a.N = N; a.num_k_iters = (K + BK - 1) / BK; a.m_tiles = (M + BM - 1) / BM;
  a.conv = 0; a.out = out; a.ldc = ldc; a.out_fp32 = out_fp32; a.relu = relu; a.bias = bias;
  a.gn_sums = gn_sums; a.gn_G = gn_G; a.gn_cpg = gn_sums ? N / gn_G : 0; a.gn_hw = gn_hw;
  return dispatch_kmajor(tmA, Bw, N, K, ldb, a, st);
}
}  // namespace lsn

// lsnet_gemm_bf16 with the full epilogue: v = acc + bias + resid[row, :];  ReLU;  v = mask[row, :] > 0 ? v : 0
extern "C" int lsnet_gemm_ex_bf16(const void* A, long long lda, const void* Bw, long long ldb, void* out, long long ldc,
                                  int M, int N, int K, const float* bias, const void* resid, long long ldr,
                                  const void* mask, long long ldm, int relu, int out_fp32, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(bias) % 16) return set_error("lsnet_gemm_ex_bf16: bias must be 16-byte aligned");
  if ((N % 16) || (K % 8) || (lda % 8) || (ldb % 8) || (ldc % (out_fp32 ? 4 : 8)) || (ldr % 8) || (ldm % 8))
    return set_error("lsnet_gemm_ex_bf16: need N%%16==0, K%%8==0 and 16-byte aligned pitches (N=%d K=%d)", N, K);
  CUtensorMap tmA;
  if (int rc = make_map_2d(&tmA, A, M, K, lda, 64, BM)) return rc;
  GemmArgs a{};
  a.M = M; a.N = N; a.num_k_iters = (K + BK - 1) / BK; a.m_tiles = (M + BM - 1) / BM;
  a.conv = 0; a.out = out; a.ldc = ldc; a.out_fp32 = out_fp32; a.relu = relu; a.bias = bias;
  a.resid = static_cast<const __nv_bfloat16*>(resid); a.ldr = ldr;
  a.mask = static_cast<const __nv_bfloat16*>(mask); a.ldm = ldm;
  return dispatch_kmajor(tmA, Bw, N, K, ldb, a, static_cast<cudaStream_t>(stream));
}

// out[M, N] (bf16) with N = nblk_rows * 64: column tile nt = A[:, (nt % cblks) * 64 .. +63] . Bw[nt * 64 .. +63, 0..63]^T
extern "C" int lsnet_gemm_blockdiag_bf16(const void* A, long long lda, const void* Bw, void* out, long long ldc, int M,
                                         int N, int cblks, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if ((N % 64) || cblks < 1 || (lda % 8) || (ldc % 8) || lda < 64LL * cblks)
    return set_error("lsnet_gemm_blockdiag_bf16: need N %% 64 == 0, cblks >= 1, 16-byte aligned pitches");
  CUtensorMap tmA, tmB;
  if (int rc = make_map_2d(&tmA, A, M, 64ull * cblks, lda, 64, BM)) return rc;
  if (int rc = make_map_2d(&tmB, Bw, N, 64, 64, 64, 64)) return rc;
  GemmArgs a{};
  a.M = M; a.N = N; a.num_k_iters = 1; a.m_tiles = (M + BM - 1) / BM; a.n_tiles = N / 64;
  a.conv = 0; a.out = out; a.ldc = ldc; a.out_fp32 = 0; a.relu = 0; a.bias = nullptr; a.blockdiag = cblks;
  return launch_kmajor<64, 1>(tmA, tmB, a, static_cast<cudaStream_t>(stream));
}

// The generalised implicit-GEMM correlation behind every convolution entry point:
//   out[b, ho*osh + ooh, wo*osw + oow, n] = sum_t sum_c x[b, ho*ish + tap_dy[t], wo*isw + tap_dx[t], c] * Wt[n, t*Cpad + c]
//                                           (+ bias[n]) (+ resid) (ReLU) (zeroed where mask <= 0),   ho < Ho, wo < Wo
extern "C" int lsnet_conv2d_taps_bf16(const void* x, int B, int H, int W, int C, long long ldp, const void* Wt, int N,
                                      int ntaps, const int* tap_dy, const int* tap_dx, const int* tap_kblk, int wt_taps,
                                      int ish, int isw, int Ho, int Wo,
                                      void* out, long long ldc, int osh, int osw, int ooh, int oow, int OH, int OW,
                                      const float* bias, const void* resid, long long ldr, const void* mask,
                                      long long ldm, int relu, int out_fp32, void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if (reinterpret_cast<uintptr_t>(bias) % 16) return set_error("lsnet_conv2d_taps_bf16: bias must be 16-byte aligned");
  if ((C % 8) || (N % 16) || (ldp % 8) || (ldc % (out_fp32 ? 4 : 8)) || (ldr % 8) || (ldm % 8))
    return set_error("lsnet_conv2d_taps_bf16: need C%%8==0, N%%16==0, aligned pitches (C=%d N=%d)", C, N);
  if (wt_taps < 1 || wt_taps > 127) return set_error("lsnet_conv2d_taps_bf16: weight pack of 1..127 taps");
  if (ntaps < 1 || ntaps > kMaxTaps || ish < 1 || isw < 1 || ish > 8 || isw > 8 || osh < 1 || osw < 1)
    return set_error("lsnet_conv2d_taps_bf16: 1..%d taps, input strides 1..8 (ntaps=%d ish=%d isw=%d)", kMaxTaps, ntaps, ish, isw);
  const int cblks = (C + 63) / 64;   // weights are packed with C padded to 64 per tap; A's channel tail is TMA zero-fill
  int TH = 8, TW = 16;
  pick_patch(Ho, Wo, BM, &TH, &TW);
  while (TW * isw > 256) { TW /= 2; TH *= 2; }
  while (TH * ish > 256) { TH /= 2; TW *= 2; }
  CUtensorMap tmA;
  if (int rc = make_map_nhwc(&tmA, x, B, H, W, C, ldp, TW, TH, isw, ish)) return rc;
  GemmArgs a{};
  a.conv = 1; a.Ho = Ho; a.Wo = Wo; a.TH = TH; a.TW = TW;
  a.tw_shift = 0;
  while ((1 << a.tw_shift) < TW) ++a.tw_shift;
  a.tiles_h = (Ho + TH - 1) / TH; a.tiles_w = (Wo + TW - 1) / TW;
  a.cblks = cblks; a.ish = ish; a.isw = isw; a.osh = osh; a.osw = osw; a.ooh = ooh; a.oow = oow; a.OH = OH; a.OW = OW;
  for (int t = 0; t < ntaps; ++t) {
    if (tap_dy[t] < -128 || tap_dy[t] > 127 || tap_dx[t] < -128 || tap_dx[t] > 127)
      return set_error("lsnet_conv2d_taps_bf16: tap offset out of range");
    a.tap_dy[t] = static_cast<signed char>(tap_dy[t]);
    a.tap_dx[t] = static_cast<signed char>(tap_dx[t]);
    const int kb = tap_kblk ? tap_kblk[t] : t;
    if (kb < 0 || kb >= wt_taps) return set_error("lsnet_conv2d_taps_bf16: weight block %d outside the %d-tap pack", kb, wt_taps);
    a.tap_kb[t] = static_cast<signed char>(kb);
  }
  a.M = B * Ho * Wo; a.N = N; a.num_k_iters = ntaps * a.cblks; a.m_tiles = B * a.tiles_h * a.tiles_w;
  a.out = out; a.ldc = ldc; a.out_fp32 = out_fp32; a.relu = relu; a.bias = bias;
  a.resid = static_cast<const __nv_bfloat16*>(resid); a.ldr = ldr;
  a.mask = static_cast<const __nv_bfloat16*>(mask); a.ldm = ldm;
  return dispatch_kmajor(tmA, Wt, N, wt_taps * cblks * 64, static_cast<long long>(wt_taps) * cblks * 64, a,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int lsnet_conv2d_nhwc_bf16(const void* x, int B, int H, int W, int C, long long ldp, const void* Wt,
                                      int N, int kh, int kw, int pad_h, int pad_w, int dil_h, int dil_w, void* out,
                                      long long ldc, const float* bias, int relu, int out_fp32, void* stream) {
  // "same" geometry only: output grid == input grid (stride 1, 2*pad == dil*(k-1))
  if (2 * pad_h != dil_h * (kh - 1) || 2 * pad_w != dil_w * (kw - 1))
    return set_error("lsnet_conv2d_nhwc_bf16: only stride-1 'same' convolutions are supported");
  if (kh * kw > kMaxTaps) return set_error("lsnet_conv2d_nhwc_bf16: at most %d taps", kMaxTaps);
  int tdy[kMaxTaps], tdx[kMaxTaps];
  for (int t = 0; t < kh * kw; ++t) { tdy[t] = (t / kw) * dil_h - pad_h; tdx[t] = (t % kw) * dil_w - pad_w; }
  return lsnet_conv2d_taps_bf16(x, B, H, W, C, ldp, Wt, N, kh * kw, tdy, tdx, nullptr, kh * kw, 1, 1, H, W, out, ldc, 1, 1, 0, 0, H, W,
                                bias, nullptr, 0, nullptr, 0, relu, out_fp32, stream);
}

// Split-K factor of a pixel-reduction GEMM with `base` output tiles over `k_chunks` 64-pixel chunks: the persistent grid
// runs ceil(base * s / SMs) rounds of ceil(k_chunks / s) chunks (+ an epilogue worth ~8 chunks of reds per round), so the
// factor must be chosen on that product -- rounding the SM count UP to a multiple of `base` (r01: 18 tiles -> 9 splits ->
// 162 items on 148 SMs) leaves a second round for a handful of CTAs and nearly doubles the kernel's time.
int lsn::pick_splits(int base, int k_chunks) {
  const int sms = num_sms();
  long long best_cost = -1;
  int best = 1;
  const int smax = k_chunks < 4 * sms ? k_chunks : 4 * sms;
  for (int s = 1; s <= smax; ++s) {
    const long long rounds = (static_cast<long long>(base) * s + sms - 1) / sms;
    const long long cost = rounds * ((k_chunks + s - 1) / s + 8);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

template <int MT>
static int launch_wgrad_t(const CUtensorMap& tmA, const CUtensorMap& tmB, WgradArgs& a, cudaStream_t st) {
  using Cfg = WCfg<MT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mnmajor_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error("cudaFuncSetAttribute(gemm_mnmajor<%d>): %s", MT, cudaGetErrorString(e));
    attr_done = true;
  }
  const int m_groups = (a.m_tiles + MT - 1) / MT;
  int base = a.taps * m_groups * a.n_tiles;
  a.splits = pick_splits(base, a.k_chunks);
  int splits = a.splits;
  int items = base * splits;
  int grid = items < num_sms() ? items : num_sms();
  const int th = timing_begin(TC_WGRAD, 2.0 * a.M * a.N * a.taps * a.k_chunks * BK, st);
  gemm_mnmajor_kernel<MT><<<grid, kThreads, Cfg::kSmemBytes, st>>>(tmA, tmB, a);
  timing_end(th, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm_mnmajor<%d> launch: %s", MT, cudaGetErrorString(e));
  count_launch();
  return 0;
}

static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, WgradArgs& a, cudaStream_t st) {
  // MT = 2 (256-row accumulator, B tile shared by both halves) measured no faster than MT = 1 with a double-buffered
  // accumulator on the LSNet shapes (r01: DCN dW 0.130 vs 0.126 ms, whole step +0.4 ms), so it is opt-in.
  static int mt2 = -1;
  if (mt2 < 0) { const char* e = getenv("LSNET_WGRAD_MT2"); mt2 = (e && e[0] == '1') ? 1 : 0; }
  return (mt2 && a.m_tiles >= 2) ? launch_wgrad_t<2>(tmA, tmB, a, st) : launch_wgrad_t<1>(tmA, tmB, a, st);
}

// out[M,N] (fp32, pre-zeroed by the caller or accumulated into) += A[P,M]^T . B[P,N]
extern "C" int lsnet_gemm_tn_bf16(const void* A, long long lda, const void* Bm, long long ldb, float* out,
                                  long long ldc, int P, int M, int N, void* stream) {
  if (P <= 0 || M <= 0 || N <= 0) return 0;
  if ((M % 8) || (N % 8) || (lda % 8) || (ldb % 8) || (ldc % 4))
    return set_error("lsnet_gemm_tn_bf16: need M%%8==0, N%%8==0 and aligned pitches");
  CUtensorMap tmA, tmB;
  if (int rc = make_map_2d(&tmA, A, P, M, lda, 64, 64)) return rc;
  if (int rc = make_map_2d(&tmB, Bm, P, N, ldb, 64, 64)) return rc;
  WgradArgs a{};
  a.M = M; a.N = N; a.m_tiles = (M + BM - 1) / BM; a.n_tiles = (N + BNW - 1) / BNW; a.taps = 1;
  a.k_chunks = (P + BK - 1) / BK; a.conv = 0; a.out = out; a.ldc = ldc;
  return launch_wgrad(tmA, tmB, a, static_cast<cudaStream_t>(stream));
}

// dW[N_out, kh*kw, C] (fp32) += sum_{output pixels p} dY[p, n] * X[p*stride - pad + tap*dil, c]   (dY on the Ho x Wo grid)
extern "C" int lsnet_conv2d_wgrad_strided_nhwc_bf16(const void* dy, long long ldy, const void* x, long long ldx, int B,
                                                    int H, int W, int C, int Ho, int Wo, int N, int kh, int kw,
                                                    int stride_h, int stride_w, int pad_h, int pad_w, int dil_h,
                                                    int dil_w, float* dw, void* stream) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  if ((C % 8) || (N % 8) || (ldy % 8) || (ldx % 8)) return set_error("lsnet_conv2d_wgrad_nhwc_bf16: alignment");
  if (stride_h < 1 || stride_w < 1 || stride_h > 8 || stride_w > 8) return set_error("lsnet_conv2d_wgrad_nhwc_bf16: stride 1..8");
  int TH = 4, TW = 16;
  pick_patch(Ho, Wo, BK, &TH, &TW);
  CUtensorMap tmA, tmB;
  if (int rc = make_map_nhwc(&tmA, dy, B, Ho, Wo, N, ldy, TW, TH, 1, 1)) return rc;
  if (int rc = make_map_nhwc(&tmB, x, B, H, W, C, ldx, TW, TH, stride_w, stride_h)) return rc;
  WgradArgs a{};
  a.M = N; a.N = C; a.m_tiles = (N + BM - 1) / BM; a.n_tiles = (C + BNW - 1) / BNW; a.taps = kh * kw;
  a.conv = 1; a.TH = TH; a.TW = TW; a.tiles_h = (Ho + TH - 1) / TH; a.tiles_w = (Wo + TW - 1) / TW;
  a.k_chunks = B * a.tiles_h * a.tiles_w; a.kw = kw; a.pad_h = pad_h; a.pad_w = pad_w; a.dil_h = dil_h;
  a.dil_w = dil_w; a.ish = stride_h; a.isw = stride_w; a.out = dw; a.ldc = static_cast<long long>(kh) * kw * C;
  return launch_wgrad(tmA, tmB, a, static_cast<cudaStream_t>(stream));
}

// stride-1 'same' conv weight gradient
extern "C" int lsnet_conv2d_wgrad_nhwc_bf16(const void* dy, long long ldy, const void* x, long long ldx, int B, int H,
                                            int W, int C, int N, int kh, int kw, int pad_h, int pad_w, int dil_h,
                                            int dil_w, float* dw, void* stream) {
  return lsnet_conv2d_wgrad_strided_nhwc_bf16(dy, ldy, x, ldx, B, H, W, C, H, W, N, kh, kw, 1, 1, pad_h, pad_w, dil_h,
                                              dil_w, dw, stream);
}
