// Internal helpers shared by the translation units of liblsnet_sm100.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lsn {
// printf-style; stores the message for lsnet_last_error() and returns a non-zero status.
int set_error(const char* fmt, ...);
// every kernel launch of this library bumps the counter reported by lsnet_launch_count()
void count_launch();
int check_launch(const char* what);
// kernel classes for the optional device timing (api.cu)
// TC_GEMM_HBM: gemm_kmajor launches whose arithmetic intensity (FLOPs / algorithmic bytes) is below the machine's ridge
// (measured 1346.8 TFLOP/s / 6.54 TB/s ~ 200 FLOP/B): bounded by HBM, so their work is accounted in bytes
enum { TC_GEMM = 0, TC_WGRAD = 1, TC_IM2COL = 2, TC_COL2IM = 3, TC_DCN_FWD = 4, TC_DCN_WGRAD = 5, TC_DCN_BWD = 6,
       TC_GEMM_HBM = 7, TC_NUM = 8 };
int timing_begin(int cls, double work, cudaStream_t st);
void timing_end(int handle, cudaStream_t st);
// host helpers of the tcgen05 GEMM family (gemm_tcgen05.cu)
// rank-2 bf16 SWIZZLE_128B map over a row-major [rows, cols] matrix with row pitch ld (elements); box = [box_cols, box_rows]
int make_map_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                uint32_t box_rows);
// rank-4 bf16 SWIZZLE_128B map over NHWC [B,H,W,C] (pixel pitch ldp elements); box = [64, TW, TH, 1]
int make_map_nhwc(CUtensorMap* m, const void* ptr, uint64_t B, uint64_t H, uint64_t W, uint64_t C, uint64_t ldp,
                  uint32_t TW, uint32_t TH, uint32_t sw = 1, uint32_t sh = 1);
// split-K factor of a pixel-reduction GEMM: `base` output tiles, `k_chunks` 64-pixel chunks (cost model over SM rounds)
int pick_splits(int base, int k_chunks);
// rank-5 map over a [B*Ho*Wo, taps*C] bf16 matrix viewed as [B, Ho, Wo, taps, C]; box = [64, taps, PW, PH, 1]
int make_map_col5d(CUtensorMap* m, const void* ptr, uint64_t B, uint64_t Ho, uint64_t Wo, uint64_t taps, uint64_t C,
                   uint64_t ldcol, uint32_t PW, uint32_t PH);
int num_sms();
// spatial patch TH x TW = pixels with the least padding waste over an H x W map
void pick_patch(int H, int W, int pixels, int* TH, int* TW);
}  // namespace lsn
