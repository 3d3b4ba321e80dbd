#include "common.cuh"
#include <cstdarg>
#include <mutex>

namespace cfun {
unsigned long long g_launches = 0;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int num_sms() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
static bool g_timing = false;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static bool g_timed = false;
void timing_begin(cudaStream_t st) {
  if (g_timing) cudaEventRecord(g_ev0, st);
}
void timing_end(cudaStream_t st) {
  if (g_timing) { cudaEventRecord(g_ev1, st); g_timed = true; }
}
}  // namespace cfun

extern "C" int cfun_kernel_timing(int on) {
  using namespace cfun;
  if (on && !g_ev0) {
    CFUN_CUDA(cudaEventCreate(&g_ev0));
    CFUN_CUDA(cudaEventCreate(&g_ev1));
  }
  g_timing = on != 0;
  g_timed = false;
  return CFUN_OK;
}
extern "C" int cfun_last_kernel_ms(float* ms) {
  using namespace cfun;
  CFUN_CHECK_ARG(ms != nullptr);
  if (!g_timed) { set_error("cfun_last_kernel_ms: no timed kernel launch since cfun_kernel_timing(1)"); return CFUN_ERR_INVALID; }
  CFUN_CUDA(cudaEventSynchronize(g_ev1));
  CFUN_CUDA(cudaEventElapsedTime(ms, g_ev0, g_ev1));
  return CFUN_OK;
}

extern "C" const char* cfun_last_error(void) { return cfun::g_err; }
extern "C" int cfun_version(void) { return 100; }
extern "C" unsigned long long cfun_launch_count(void) { return cfun::g_launches; }
extern "C" int cfun_device_is_sm100(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

// bring-up aid (CFUN_DEBUG_CAPTURE=1 in ops.py): 0 = stream not capturing, 1 = capturing, 2 = capture invalidated
extern "C" int cfun_stream_capture_status(void* stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaError_t e = cudaStreamIsCapturing(reinterpret_cast<cudaStream_t>(stream), &st);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return 2; }
  return st == cudaStreamCaptureStatusNone ? 0 : (st == cudaStreamCaptureStatusActive ? 1 : 2);
}
