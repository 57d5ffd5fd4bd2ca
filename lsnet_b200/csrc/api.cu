// C-ABI plumbing of liblsnet_sm100.so: status / error string, launch accounting, device probe.
#include <stdarg.h>
#include <stdio.h>

#include <vector>

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

// ---- optional per-kernel-class device timing (bench.py roofline): CUDA events on the launch stream ---------------
namespace lsn {
struct TimedLaunch { int cls; cudaEvent_t a, b; double work; };
static bool g_timing = false;
static std::vector<TimedLaunch> g_timed;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
bool timing_on() { return g_timing; }
// Inside a stream capture the records become EXTERNAL event-record nodes of the graph: every replay re-stamps them, so
// after a replay the elapsed time of each captured launch can be read like that of an eager launch.
static void record(cudaEvent_t e, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  if (cs == cudaStreamCaptureStatusActive) cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
  else cudaEventRecord(e, st);
}
int timing_begin(int cls, double work, cudaStream_t st) {
  if (!g_timing) return -1;
  TimedLaunch t{cls, get_event(), get_event(), work};
  record(t.a, st);
  g_timed.push_back(t);
  return static_cast<int>(g_timed.size()) - 1;
}
void timing_end(int h, cudaStream_t st) {
  if (h >= 0) record(g_timed[h].b, st);
}
}  // namespace lsn

extern "C" void lsnet_timing_enable(int on) { lsn::g_timing = on != 0; }
// Sums elapsed ms / launches / algorithmic work (FLOPs or bytes) of one kernel class since the last reset; syncs.
extern "C" int lsnet_timing_collect(int cls, double* total_ms, long long* launches, double* work) {
  *total_ms = 0; *launches = 0; *work = 0;
  for (auto& t : lsn::g_timed) {
    if (t.cls != cls) continue;
    if (cudaEventSynchronize(t.b) != cudaSuccess) return lsn::set_error("timing: event sync failed");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.a, t.b);
    *total_ms += ms; *launches += 1; *work += t.work;
  }
  return 0;
}
extern "C" void lsnet_timing_reset(void) {
  for (auto& t : lsn::g_timed) { lsn::g_pool.push_back(t.a); lsn::g_pool.push_back(t.b); }
  lsn::g_timed.clear();
}

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
