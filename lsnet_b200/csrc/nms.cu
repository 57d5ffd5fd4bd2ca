// Greedy non-maximum suppression on the device (inference decode of LSHead.get_bboxes: multiclass_nms_lsvr ->
// batched_nms -> nms, mmdet/core/post_processing/bbox_nms.py:60-99, mmdet/ops/nms/nms_wrapper.py:7-157; the reference's
// own kernel + host sweep: mmdet/ops/nms/src/cuda/nms_kernel.cu).  Boxes arrive sorted by descending score.
//   nms_mask   one CTA per 64 x 64 tile of the upper triangle: bit j of mask[i][tile] = IoU(box i, box 64*tile + j) > thr
//   nms_sweep  ONE CTA walks the 64-box blocks in order: a single thread resolves a block against the running
//              "removed" bitmap and its own diagonal word, then all threads OR the mask rows of the boxes it kept into the
//              bitmap -- no round trip to the host, the kept indices and their count stay on the device.
// IoU as the reference computes it (areas without the +1 of the legacy convention): inter / (Sa + Sb - inter).
#include "common.cuh"
#include "lsnet_internal.h"

namespace lsn {

__device__ __forceinline__ float nms_iou(const float4& a, const float4& b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
  const float inter = w * h;
  const float sa = (a.z - a.x) * (a.w - a.y), sb = (b.z - b.x) * (b.w - b.y);
  return inter / (sa + sb - inter);
}

__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, int n, float thr, unsigned long long* __restrict__ mask, int nblk) {
  const int row_blk = blockIdx.y, col_blk = blockIdx.x;
  if (col_blk < row_blk) return;                       // only later boxes can be suppressed by an earlier one
  __shared__ float4 cb[64];
  const int t = threadIdx.x;
  const int col = col_blk * 64 + t;
  if (col < n) cb[t] = boxes[col];
  __syncthreads();
  const int row = row_blk * 64 + t;
  if (row >= n) return;
  const float4 me = boxes[row];
  const int ncol = min(64, n - col_blk * 64);
  unsigned long long bits = 0;
  for (int j = (row_blk == col_blk) ? t + 1 : 0; j < ncol; ++j)
    if (nms_iou(me, cb[j]) > thr) bits |= 1ull << j;
  mask[static_cast<long long>(row) * nblk + col_blk] = bits;
}

__global__ void __launch_bounds__(256)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int n, int nblk, int* __restrict__ keep,
                 int* __restrict__ num_keep) {
  extern __shared__ unsigned long long remv[];          // [nblk] boxes already suppressed
  __shared__ unsigned long long kept_word;
  __shared__ int count;
  for (int w = threadIdx.x; w < nblk; w += blockDim.x) remv[w] = 0;
  if (threadIdx.x == 0) count = 0;
  __syncthreads();
  for (int b = 0; b < nblk; ++b) {
    if (threadIdx.x == 0) {
      unsigned long long word = remv[b], kept = 0;
      const int nb = min(64, n - b * 64);
      int c = count;
      for (int i = 0; i < nb; ++i) {
        if (!((word >> i) & 1ull)) {
          kept |= 1ull << i;
          keep[c++] = b * 64 + i;
          word |= mask[static_cast<long long>(b * 64 + i) * nblk + b];
        }
      }
      kept_word = kept;
      count = c;
    }
    __syncthreads();
    unsigned long long k = kept_word;
    while (k) {
      const int i = __ffsll(static_cast<long long>(k)) - 1;
      k &= k - 1;
      const unsigned long long* row = mask + static_cast<long long>(b * 64 + i) * nblk;
      for (int w = b + 1 + threadIdx.x; w < nblk; w += blockDim.x) remv[w] |= row[w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_keep = count;
}

}  // namespace lsn

using namespace lsn;

extern "C" size_t lsnet_nms_workspace_size(int n) {
  const long long nblk = (n + 63) / 64;
  return static_cast<size_t>(n > 0 ? static_cast<long long>(n) * nblk * 8 : 0);
}

extern "C" int lsnet_nms(const float* boxes, int n, float iou_thr, void* workspace, size_t workspace_bytes, int* keep,
                         int* num_keep, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 0) {
    cudaMemsetAsync(num_keep, 0, sizeof(int), st);
    return 0;
  }
  const int nblk = (n + 63) / 64;
  if (!workspace || workspace_bytes < lsnet_nms_workspace_size(n))
    return set_error("lsnet_nms: needs %zu workspace bytes (got %zu)", lsnet_nms_workspace_size(n), workspace_bytes);
  if (nblk > 65535) return set_error("lsnet_nms: %d boxes exceed the supported 4 194 240", n);
  if (static_cast<size_t>(nblk) * 8 > 200 * 1024) return set_error("lsnet_nms: %d boxes exceed the sweep's shared memory", n);
  unsigned long long* mask = static_cast<unsigned long long*>(workspace);
  nms_mask_kernel<<<dim3(nblk, nblk), 64, 0, st>>>(reinterpret_cast<const float4*>(boxes), n, iou_thr, mask, nblk);
  if (int rc = check_launch("nms_mask")) return rc;
  const size_t smem = static_cast<size_t>(nblk) * 8;
  if (smem > 48 * 1024) cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  nms_sweep_kernel<<<1, 256, smem, st>>>(mask, n, nblk, keep, num_keep);
  return check_launch("nms_sweep");
}
