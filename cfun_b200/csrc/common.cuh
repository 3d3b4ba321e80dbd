// Shared helpers for libcfun_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../../include/cfun_b200.h"

namespace cfun {

void set_error(const char* fmt, ...);

#define CFUN_CHECK_ARG(cond)                                                         \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      cfun::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond);     \
      return CFUN_ERR_INVALID;                                                       \
    }                                                                                \
  } while (0)

#define CFUN_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      cfun::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));     \
      return CFUN_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

extern unsigned long long g_launches;  // kernels launched by this library (bench.py "gpu_launches")

#define CFUN_LAUNCH_CHECK()                                                                      \
  do {                                                                                           \
    ++cfun::g_launches;                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                         \
    if (e_ != cudaSuccess) {                                                                     \
      cfun::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return CFUN_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();

// Measurement hook (cfun_kernel_timing / cfun_last_kernel_ms): when switched on, the tensor-core conv launchers bracket their
// MAIN kernel (not the operand packs) with CUDA events on the launching stream, so that bench.py can report the roofline of
// the kernel itself, live, without a profiler.  Off by default; costs nothing then.
void timing_begin(cudaStream_t st);
void timing_end(cudaStream_t st);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cfun
