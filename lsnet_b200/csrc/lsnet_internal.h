// Internal helpers shared by the translation units of liblsnet_sm100.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace lsn {
// printf-style; stores the message for lsnet_last_error() and returns a non-zero status.
int set_error(const char* fmt, ...);
// every kernel launch of this library bumps the counter reported by lsnet_launch_count()
void count_launch();
int check_launch(const char* what);
// kernel classes for the optional device timing (api.cu)
enum { TC_GEMM = 0, TC_WGRAD = 1, TC_IM2COL = 2, TC_COL2IM = 3 };
int timing_begin(int cls, double work, cudaStream_t st);
void timing_end(int handle, cudaStream_t st);
}  // namespace lsn
