// common.cuh -- shared helpers for libcreste_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/creste_b200.h"

namespace creste {

void set_error(const char* fmt, ...);

#define CRESTE_CHECK_ARG(cond, ...)                 \
  do {                                              \
    if (!(cond)) {                                  \
      creste::set_error(__VA_ARGS__);               \
      return CRESTE_ERR_ARG;                        \
    }                                               \
  } while (0)

#define CRESTE_CUDA(call)                                                            \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      creste::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                        __FILE__, __LINE__);                                         \
      return (int)e__;                                                               \
    }                                                                                \
  } while (0)

void count_launch();

inline int launch_check(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int num_sms();

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace creste
