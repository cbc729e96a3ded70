// lidar.cu -- LiDAR sweep -> sparse depth raster (K18) and softmax-expectation depth (K4).
//
// creste_lidar_raster replaces pixels_to_depth (reference creste/utils/projection.py:64-134:
// float64 projection, truncation toward zero to int32, z>0 & in-image mask, per-pixel MAX depth
// via torch_scatter) and the millimetre quantisation of scripts/preprocessing/
// build_dense_depth.py:461-463 (float32 * 1000, clip, uint16 truncation).  Positive IEEE doubles
// order like unsigned integers, so the per-pixel max is one 64-bit atomicMax on the bit pattern.
// 131072 points, 3.5 MB algorithmic traffic: pure latency/HBM work, no tensor cores.
//
// creste_depth_expectation replaces convert_to_metric_depth_differentiable (reference
// creste/utils/depth_utils.py:300-313) and the argmax of creste/models/depth.py:70: one warp per
// pixel over the 128 NHWC-contiguous bins (float4 per lane), warp-shuffle max / sum.
#include "common.cuh"

namespace creste {

struct P34 { double m[12]; };

__global__ void lidar_project_kernel(const float* __restrict__ pc, int npts, int stride, P34 P,
                                     int H, int W, unsigned long long* __restrict__ zbuf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  const double x = (double)pc[(size_t)i * stride], y = (double)pc[(size_t)i * stride + 1],
               z = (double)pc[(size_t)i * stride + 2];
  double c[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double acc = __dmul_rn(P.m[r * 4 + 0], x);
    acc = __fma_rn(P.m[r * 4 + 1], y, acc);
    acc = __fma_rn(P.m[r * 4 + 2], z, acc);
    acc = __fma_rn(P.m[r * 4 + 3], 1.0, acc);
    c[r] = acc;
  }
  if (!(c[2] > 0.0)) return;
  double u = __ddiv_rn(c[0], c[2]), v = __ddiv_rn(c[1], c[2]);
  if (u != u || v != v) return;
  u = fmin(fmax(u, -2147483648.0), 2147483647.0);
  v = fmin(fmax(v, -2147483648.0), 2147483647.0);
  const long long ui = (long long)u, vi = (long long)v;  // truncation toward zero (astype(int32))
  if (ui < 0 || ui >= W || vi < 0 || vi >= H) return;
  atomicMax(zbuf + (size_t)vi * W + ui, (unsigned long long)__double_as_longlong(c[2]));
}

__global__ void lidar_finish_kernel(const unsigned long long* __restrict__ zbuf, int n,
                                    float* __restrict__ depth_m, float* __restrict__ depth_mm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = (float)__longlong_as_double((long long)zbuf[i]);
  if (depth_m) depth_m[i] = d;
  if (depth_mm) {
    float mm = __fmul_rn(d, 1000.0f);
    mm = fminf(fmaxf(mm, 0.0f), 65535.0f);
    depth_mm[i] = (float)(unsigned short)mm;
  }
}

// one warp per pixel, D == 128
__global__ void __launch_bounds__(256) depth_expectation_kernel(const float* __restrict__ logits,
                                                                int NP, float dmin, float dmax, float out_div,
                                                                float* __restrict__ metric,
                                                                long long* __restrict__ bins) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float step = __fdiv_rn(__fsub_rn(dmax, dmin), 127.0f);
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < NP; p += gridDim.x * wpb) {
    const float4 l4 = __ldg(reinterpret_cast<const float4*>(logits + (size_t)p * 128) + lane);
    const float l[4] = {l4.x, l4.y, l4.z, l4.w};
    // arg-max, first index on ties
    float m = l[0];
    int arg = lane * 4;
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (l[j] > m) { m = l[j]; arg = lane * 4 + j; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > m || (om == m && oa < arg)) { m = om; arg = oa; }
    }
    float s = 0.0f, e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { e[j] = expf(__fsub_rn(l[j], m)); s += e[j]; }
    s = warp_sum(s);
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = lane * 4 + j;
      // torch.linspace: start + step*i for the first half, end - step*(n-1-i) for the second
      const float val = (k < 64) ? __fmaf_rn(step, (float)k, dmin) : __fsub_rn(dmax, __fmul_rn(step, (float)(127 - k)));
      acc = __fmaf_rn(__fdiv_rn(e[j], s), val, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      metric[p] = __fdiv_rn(acc, out_div);      // 1000: metres (depth.py:100); 1: the bin units (mm)
      if (bins) bins[p] = arg;
    }
  }
}

// bin_depths (depth_utils.py:346-383): same fp32 operation order as the tensor expression
__global__ void bin_depths_kernel(const float* __restrict__ d, long long n, int mode, float dmin, float bin_size,
                                  float log_lo, float log_span, int num_bins, int target,
                                  float* __restrict__ out_f, long long* __restrict__ out_i) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = d[i];
  float idx;
  if (mode == 0) {
    idx = __fdiv_rn(__fsub_rn(x, dmin), bin_size);
  } else if (mode == 1) {
    const float t = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(8.0f, __fsub_rn(x, dmin)), bin_size));
    idx = __fadd_rn(-0.5f, __fmul_rn(0.5f, sqrtf(t)));
  } else {
    idx = __fdiv_rn(__fmul_rn((float)num_bins, __fsub_rn(logf(__fadd_rn(1.0f, x)), log_lo)), log_span);
  }
  if (!target) { out_f[i] = idx; return; }
  const bool bad = (idx < 0.0f) || (idx > (float)num_bins) || !isfinite(idx);
  out_i[i] = bad ? (long long)num_bins : (long long)idx;
}

}  // namespace creste

using namespace creste;

extern "C" int creste_bin_depths(const float* depth, long long n, int mode, float depth_min, float depth_max,
                                 int num_bins, int target, float* out_f, int64_t* out_i, void* stream) {
  CRESTE_CHECK_ARG(depth && n > 0 && num_bins > 0 && mode >= 0 && mode <= 2, "creste_bin_depths: bad args");
  CRESTE_CHECK_ARG(target ? (out_i != nullptr) : (out_f != nullptr), "creste_bin_depths: output pointer");
  // python-float (double) constants rounded once to fp32 when they meet the fp32 tensor, as in torch
  float bin_size = 1.0f, log_lo = 0.0f, log_span = 1.0f;
  if (mode == 0) bin_size = (float)(((double)depth_max - (double)depth_min) / (double)num_bins);
  if (mode == 1) bin_size = (float)(2.0 * ((double)depth_max - (double)depth_min) / ((double)num_bins * (1.0 + num_bins)));
  if (mode == 2) {
    log_lo = (float)log(1.0 + (double)depth_min);
    log_span = (float)(log(1.0 + (double)depth_max) - log(1.0 + (double)depth_min));
  }
  bin_depths_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      depth, n, mode, depth_min, bin_size, log_lo, log_span, num_bins, target, out_f, (long long*)out_i);
  return launch_check("bin_depths_kernel");
}

extern "C" int creste_lidar_raster(const float* pc, int npts, int stride, const double* P34_host,
                                   int H, int W, float* depth_m, float* depth_mm, void* ws,
                                   size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG((pc || npts == 0) && P34_host && ws, "creste_lidar_raster: null pointer");
  CRESTE_CHECK_ARG(npts >= 0 && stride >= 3 && H > 0 && W > 0, "creste_lidar_raster: bad shape");
  const size_t need = (size_t)H * W * 8;
  if (ws_bytes < need) {
    set_error("creste_lidar_raster: workspace %zu < %zu", ws_bytes, need);
    return CRESTE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  P34 P;
  for (int i = 0; i < 12; ++i) P.m[i] = P34_host[i];
  CRESTE_CUDA(cudaMemsetAsync(ws, 0, need, st));
  if (npts > 0) {
    lidar_project_kernel<<<ceil_div(npts, 256), 256, 0, st>>>(pc, npts, stride, P, H, W,
                                                              (unsigned long long*)ws);
    int rc = launch_check("lidar_project_kernel");
    if (rc) return rc;
  }
  lidar_finish_kernel<<<ceil_div(H * W, 256), 256, 0, st>>>((const unsigned long long*)ws, H * W,
                                                            depth_m, depth_mm);
  return launch_check("lidar_finish_kernel");
}

extern "C" int creste_depth_expectation(const float* logits, int NP, int D, float depth_min_mm,
                                        float depth_max_mm, float out_div, float* metric, int64_t* bins,
                                        void* stream) {
  CRESTE_CHECK_ARG(logits && metric, "creste_depth_expectation: null pointer");
  CRESTE_CHECK_ARG(D == 128, "creste_depth_expectation: only D = 128 bins is implemented (got %d)", D);
  CRESTE_CHECK_ARG(NP > 0, "creste_depth_expectation: bad shape");
  const int blocks = min(ceil_div(NP, 8), 148 * 8);
  depth_expectation_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      logits, NP, depth_min_mm, depth_max_mm, out_div, metric, (long long*)bins);
  return launch_check("depth_expectation_kernel");
}
