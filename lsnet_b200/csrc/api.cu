// C-ABI plumbing of liblsnet_sm100.so: status / error string, launch accounting, device probe.
#include <stdarg.h>
#include <stdio.h>

#include "lsnet_internal.h"

namespace lsn {
static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}
void count_launch() { ++g_launches; }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("%s launch failed: %s", what, cudaGetErrorString(e));
  ++g_launches;
  return 0;
}
}  // namespace lsn

extern "C" const char* lsnet_last_error(void) { return lsn::g_err; }
extern "C" unsigned long long lsnet_launch_count(void) { return lsn::g_launches; }
extern "C" int lsnet_abi_version(void) { return 1; }
// 0 when a CUDA device of compute capability 10.x is current; non-zero (+ error string) otherwise.
extern "C" int lsnet_require_sm100(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return lsn::set_error("no CUDA device: %s", cudaGetErrorString(e));
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return lsn::set_error("liblsnet_sm100 needs an sm_100 device, found compute capability %d.x", major);
  return 0;
}
