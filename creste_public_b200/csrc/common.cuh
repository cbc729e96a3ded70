// common.cuh -- shared helpers for libcreste_b200 (sm_100a only).
#pragma once
#include <cuda_fp16.h>
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
// 3xFP16 operand split of four values already multiplied by the power-of-two scale: hi = fp16_rn(x), lo =
// fp16_rn((x - hi) * 2^11), packed as two uint2 (channel order).  Packed conversions (cvt.rn.f16x2.f32): the same
// round-to-nearest results as four scalar conversions at half the conversion-pipe instructions and no byte shuffles.
__device__ __forceinline__ void split4_f16(float x0, float x1, float x2, float x3, uint2& hi, uint2& lo) {
  const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn((x0 - f01.x) * 2048.0f, (x1 - f01.y) * 2048.0f);
  const __half2 l23 = __floats2half2_rn((x2 - f23.x) * 2048.0f, (x3 - f23.y) * 2048.0f);
  hi = make_uint2(*reinterpret_cast<const unsigned*>(&h01), *reinterpret_cast<const unsigned*>(&h23));
  lo = make_uint2(*reinterpret_cast<const unsigned*>(&l01), *reinterpret_cast<const unsigned*>(&l23));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace creste
