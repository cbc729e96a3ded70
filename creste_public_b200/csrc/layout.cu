// layout.cu -- memory-bound glue kernels (NHWC): layout shuffles at the module boundary,
// bilinear upsample + channel concat, 2x2 max-pool + concat + crop, expert-visitation raster.
// All are pure HBM streaming kernels: float4-vectorised, channel-contiguous, grid-stride.
#include <cuda_fp16.h>

#include "common.cuh"

namespace creste {

// ---- NCHW <-> NHWC through a 32x32 shared tile (both sides coalesced)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows,
                                                        int cols, float* __restrict__ out) {
  // in: [batch][rows][cols] -> out: [batch][cols][rows]
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < rows && c < cols) ? in[base + (size_t)r * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < rows && c < cols) out[base + (size_t)c * rows + r] = tile[tx][j];
  }
}

// ---- projection head: 1x1 conv with a handful of output channels (the DeconvHead `proj`, K = 32 / 6 / 2 over
// C = 128 features at 256 x 256) fused with the layout change of its INPUT.  The generic conv kernels tile the
// output channels by >= 32 and waste 5-16x of their FFMA work here, and the reference's output dict wants both the
// predictions and the 128-channel features in NCHW -- another full read + write of the feature tensor.  One pass:
// a warp stages 32 pixels x 32 channels through a padded shared tile (coalesced NHWC read), lane = pixel then
// accumulates its K dot products in registers (weights broadcast from shared memory, channels in ascending order:
// the same fp32 FFMA chain as conv_simt_kernel) and writes the tile back transposed as 32 coalesced NCHW rows.
// HBM traffic: read x once, write x^T once, write the predictions -- the floor for the two outputs.
template <int KO>
__global__ void __launch_bounds__(256) proj_head_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, long long M, int C,
                                                        long long PQ, float* __restrict__ pred_nhwc,
                                                        float* __restrict__ pred_nchw, float* __restrict__ x_nchw,
                                                        int K) {
  extern __shared__ float s_w[];                       // [KO][C]
  __shared__ float tile[8][32][33];
  for (int i = threadIdx.x; i < KO * C; i += blockDim.x) s_w[i] = (i / C < K) ? w[i] : 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long p0 = ((long long)blockIdx.x * 8 + warp) * 32; p0 < M; p0 += (long long)gridDim.x * 256) {
    const long long pix = p0 + lane;                   // this lane's pixel in the compute / store phases
    const bool ok = pix < M;
    const long long img = ok ? pix / PQ : 0, pp = ok ? pix - img * PQ : 0;
    float acc[KO];
#pragma unroll
    for (int k = 0; k < KO; ++k) acc[k] = 0.0f;
    for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r)                     // row r of the tile = pixel p0 + r: one 128-byte line
        tile[warp][r][lane] = (p0 + r < M) ? __ldg(x + (size_t)(p0 + r) * C + c0 + lane) : 0.0f;
      __syncwarp();
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        const float xv = tile[warp][lane][c];
#pragma unroll
        for (int k = 0; k < KO; ++k) acc[k] = fmaf(xv, s_w[k * C + c0 + c], acc[k]);
        if (x_nchw && ok) x_nchw[((size_t)img * C + c0 + c) * PQ + pp] = xv;
      }
      __syncwarp();
    }
    if (ok) {
#pragma unroll
      for (int k = 0; k < KO; ++k) {
        if (k < K) {
          const float v = acc[k] + (bias ? __ldg(bias + k) : 0.0f);
          if (pred_nhwc) pred_nhwc[(size_t)pix * K + k] = v;
          if (pred_nchw) pred_nchw[((size_t)img * K + k) * PQ + pp] = v;
        }
      }
    }
  }
}

// ---- cat([skip, bilinear(x)], C) in NHWC; Cs % 4 == 0, Cx % 4 == 0.
// PyTorch upsample_bilinear2d (align_corners=False): src = max(r*(dst+0.5)-0.5, 0); i0 = (int)src;
// i1 = i0 + (i0 < in-1); l1 = src - i0; l0 = 1 - l1;
// out = l0h*(l0w*v00 + l1w*v01) + l1h*(l0w*v10 + l1w*v11).
// SPLIT: instead of the fp32 tensor, write the 3xFP16 operand of the consuming tensor-core conv directly -- hi =
// fp16(v * s), lo = fp16((v * s - hi) * 2^11) with the power-of-two scale s taken from the amax bounds that travel
// with the two inputs (an interpolation / concat never exceeds the maximum of its inputs).  Saves the fp32 write, the
// split pre-pass's read of it and one launch per consumer; same arithmetic per element as f16_split_kernel.
template <bool SPLIT>
__global__ void __launch_bounds__(256) upsample_concat_kernel(const float* __restrict__ skip, int Cs,
                                                              const float* __restrict__ x, int N,
                                                              int Hi, int Wi, int Cx, int Ho, int Wo,
                                                              float rh, float rw, int x_first,
                                                              float* __restrict__ out,
                                                              const unsigned* __restrict__ amax_a,
                                                              const unsigned* __restrict__ amax_b,
                                                              uint2* __restrict__ hi, uint2* __restrict__ lo,
                                                              float* __restrict__ scal) {
  float sc = 1.0f;
  if (SPLIT) {
    unsigned b = __ldg(amax_a);
    if (amax_b) b = max(b, __ldg(amax_b));
    int e = (int)((b >> 23) & 0xffu) - 127;
    if (b == 0u || !isfinite(__uint_as_float(b))) e = 14;
    const int k = max(-100, min(100, 14 - e));
    sc = __uint_as_float((unsigned)(127 + k) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal[0] = sc; scal[1] = __uint_as_float((unsigned)(127 - k) << 23); }
  }
  const int Ct = Cs + Cx;
  const int c4t = Ct / 4, cs4 = Cs / 4, cx4 = Cx / 4;
  // one warp per output pixel, lanes over the 4-channel groups: the pixel decode (three integer divisions), the
  // source taps and the interpolation weights are computed once per pixel per warp instead of once per element
  // (five 64-bit divisions per float4 made the element-per-thread form issue-bound at ~30 % of HBM rate)
  const int lane = threadIdx.x & 31;
  const long long npix = (long long)N * Ho * Wo;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pix = warp0; pix < npix; pix += nwarps) {
    const int ox = (int)(pix % Wo);
    const long long t = pix / Wo;
    const int oy = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const float sy = fmaxf(__fsub_rn(__fmul_rn(rh, __fadd_rn((float)oy, 0.5f)), 0.5f), 0.0f);
    const float sx = fmaxf(__fsub_rn(__fmul_rn(rw, __fadd_rn((float)ox, 0.5f)), 0.5f), 0.0f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float l1h = __fsub_rn(sy, (float)y0), l0h = __fsub_rn(1.0f, l1h);
    const float l1w = __fsub_rn(sx, (float)x0), l0w = __fsub_rn(1.0f, l1w);
    const float* b = x + (size_t)n * Hi * Wi * Cx;
    const float4* p00 = reinterpret_cast<const float4*>(b + ((size_t)y0 * Wi + x0) * Cx);
    const float4* p01 = reinterpret_cast<const float4*>(b + ((size_t)y0 * Wi + x1) * Cx);
    const float4* p10 = reinterpret_cast<const float4*>(b + ((size_t)y1 * Wi + x0) * Cx);
    const float4* p11 = reinterpret_cast<const float4*>(b + ((size_t)y1 * Wi + x1) * Cx);
    const float4* ps = Cs ? reinterpret_cast<const float4*>(skip + pix * Cs) : nullptr;
    for (int c4 = lane; c4 < c4t; c4 += 32) {
      float4 v;
      const bool is_skip = x_first ? (c4 >= cx4) : (c4 < cs4);
      if (is_skip) {
        v = __ldg(ps + (x_first ? c4 - cx4 : c4));
      } else {
        const int cc = x_first ? c4 : c4 - cs4;
        const float4 v00 = __ldg(p00 + cc), v01 = __ldg(p01 + cc), v10 = __ldg(p10 + cc), v11 = __ldg(p11 + cc);
#define CRESTE_LERP(f) (l0h * (l0w * v00.f + l1w * v01.f) + l1h * (l0w * v10.f + l1w * v11.f))
        v.x = CRESTE_LERP(x); v.y = CRESTE_LERP(y); v.z = CRESTE_LERP(z); v.w = CRESTE_LERP(w);
#undef CRESTE_LERP
      }
      if (!SPLIT) {
        reinterpret_cast<float4*>(out + pix * Ct)[c4] = v;
      } else {
        const float xs[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
        unsigned short h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half hh = __float2half_rn(xs[j]);
          h[j] = __half_as_ushort(hh);
          l[j] = __half_as_ushort(__float2half_rn((xs[j] - __half2float(hh)) * 2048.0f));
        }
        hi[pix * c4t + c4] = make_uint2((unsigned)h[0] | ((unsigned)h[1] << 16), (unsigned)h[2] | ((unsigned)h[3] << 16));
        if (lo) lo[pix * c4t + c4] = make_uint2((unsigned)l[0] | ((unsigned)l[1] << 16), (unsigned)l[2] | ((unsigned)l[3] << 16));
      }
    }
  }
}

// ---- 2x2/2 max-pool of channel-concatenated NHWC sources, cropped to rows_out rows
struct PoolSrcs { const float* p[3]; int c[3]; int n; };
__global__ void __launch_bounds__(256) maxpool2_concat_kernel(PoolSrcs s, int N, int H, int W,
                                                              int rows_out, int Ct,
                                                              float* __restrict__ out_nhwc,
                                                              float* __restrict__ out_nchw) {
  const int Wo = W / 2;
  const long long total = (long long)N * rows_out * Wo * Ct;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Ct);
    const long long pix = i / Ct;
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % rows_out);
    const int n = (int)(pix / ((long long)Wo * rows_out));
    const int ctot = c;
    int k = 0;
    while (k < s.n - 1 && c >= s.c[k]) { c -= s.c[k]; ++k; }
    const float* b = s.p[k] + ((size_t)n * H * W) * s.c[k] + c;
    const int Ck = s.c[k];
    const size_t r0 = ((size_t)(2 * oy) * W + 2 * ox) * Ck, r1 = r0 + (size_t)W * Ck;
    const float m = fmaxf(fmaxf(__ldg(b + r0), __ldg(b + r0 + Ck)), fmaxf(__ldg(b + r1), __ldg(b + r1 + Ck)));
    if (out_nhwc) out_nhwc[i] = m;
    if (out_nchw) out_nchw[(((size_t)n * Ct + ctot) * rows_out + oy) * Wo + ox] = m;
  }
}

// ---- expert visitation raster (loss_utils.py:1055-1116, second definition)
template <typename T>
__global__ void expert_visitation_kernel(const T* __restrict__ traj, int B, int Tn, T map_ds,
                                         int max_steps, int H, int W, float* __restrict__ counts) {
  // one thread per (b, t, k): k-th interpolation sample of segment t; plus the appended last pose
  const long long per_b = (long long)(Tn - 1) * max_steps + 1;
  const long long total = (long long)B * per_b;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const long long j = i - (long long)b * per_b;
    T pr, pc;
    if (j == per_b - 1) {
      pr = traj[((size_t)b * Tn + Tn - 1) * 2] / map_ds;
      pc = traj[((size_t)b * Tn + Tn - 1) * 2 + 1] / map_ds;
    } else {
      const int t = (int)(j / max_steps), k = (int)(j - (long long)t * max_steps);
      const T r0 = traj[((size_t)b * Tn + t) * 2] / map_ds, c0 = traj[((size_t)b * Tn + t) * 2 + 1] / map_ds;
      const T r1 = traj[((size_t)b * Tn + t + 1) * 2] / map_ds, c1 = traj[((size_t)b * Tn + t + 1) * 2 + 1] / map_ds;
      // torch.linspace(0, 1, max_steps) is float32 whatever the trajectory dtype
      float f = 0.0f;
      if (max_steps > 1) {
        const float step = __fdiv_rn(1.0f, (float)(max_steps - 1));
        f = (k < max_steps / 2) ? __fmul_rn(step, (float)k)
                                : __fsub_rn(1.0f, __fmul_rn(step, (float)(max_steps - 1 - k)));
      }
      const T ft = (T)f;
      if (sizeof(T) == 4) {
        pr = (T)__fadd_rn((float)r0, __fmul_rn((float)ft, __fsub_rn((float)r1, (float)r0)));
        pc = (T)__fadd_rn((float)c0, __fmul_rn((float)ft, __fsub_rn((float)c1, (float)c0)));
      } else {
        pr = (T)__dadd_rn((double)r0, __dmul_rn((double)ft, __dsub_rn((double)r1, (double)r0)));
        pc = (T)__dadd_rn((double)c0, __dmul_rn((double)ft, __dsub_rn((double)c1, (double)c0)));
      }
    }
    pr = pr < (T)0 ? (T)0 : (pr > (T)(H - 1) ? (T)(H - 1) : pr);
    pc = pc < (T)0 ? (T)0 : (pc > (T)(W - 1) ? (T)(W - 1) : pc);
    const long long ir = (long long)pr, ic = (long long)pc;
    counts[((size_t)b * H + ir) * W + ic] = 1.0f;  // scatter_add of ones, then clipped to 1
  }
}

}  // namespace creste

using namespace creste;

static int grid_for(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int creste_nchw_to_nhwc(const float* in, int N, int C, int H, int W, float* out,
                                   void* stream) {
  CRESTE_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0, "creste_nchw_to_nhwc: bad args");
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, C, H * W, out);
  return launch_check("transpose_kernel");
}

extern "C" int creste_nhwc_to_nchw(const float* in, int N, int H, int W, int C, float* out,
                                   void* stream) {
  CRESTE_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0, "creste_nhwc_to_nchw: bad args");
  dim3 grid(ceil_div(C, 32), ceil_div(H * W, 32), N);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, H * W, C, out);
  return launch_check("transpose_kernel");
}

extern "C" int creste_proj_head(const float* x, const float* w, const float* bias, int N, int H, int W, int C, int K,
                                float* pred_nhwc, float* pred_nchw, float* x_nchw, void* stream) {
  CRESTE_CHECK_ARG(x && w && (pred_nhwc || pred_nchw) && N > 0 && H > 0 && W > 0, "creste_proj_head: null pointer");
  CRESTE_CHECK_ARG(C % 32 == 0 && C <= 512 && K >= 1 && K <= 32, "creste_proj_head: C %% 32 == 0, C <= 512, K <= 32");
  const long long PQ = (long long)H * W, M = (long long)N * PQ;
  long long blocks = (M + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define CRESTE_PROJ(KO)                                                                                         \
  proj_head_kernel<KO><<<(int)blocks, 256, (size_t)KO * C * sizeof(float), st>>>(x, w, bias, M, C, PQ, pred_nhwc, \
                                                                                pred_nchw, x_nchw, K)
  if (K <= 2) CRESTE_PROJ(2);
  else if (K <= 8) CRESTE_PROJ(8);
  else if (K <= 16) CRESTE_PROJ(16);
  else { 
    CRESTE_CUDA(cudaFuncSetAttribute(proj_head_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 512 * 4));
    CRESTE_PROJ(32);
  }
#undef CRESTE_PROJ
  return launch_check("proj_head_kernel");
}

extern "C" int creste_upsample_concat(const float* skip, int Cs, const float* x, int N, int Hi,
                                      int Wi, int Cx, int Ho, int Wo, float rh, float rw, int x_first,
                                      float* out, void* stream) {
  CRESTE_CHECK_ARG(x && out, "creste_upsample_concat: null pointer");
  CRESTE_CHECK_ARG((Cs == 0 || skip) && Cs % 4 == 0 && Cx % 4 == 0 && Cx > 0,
                   "creste_upsample_concat: channel counts must be multiples of 4");
  const long long total = (long long)N * Ho * Wo * 32;      // one warp per output pixel
  upsample_concat_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      skip, Cs, x, N, Hi, Wi, Cx, Ho, Wo, rh, rw, x_first, out, nullptr, nullptr, nullptr, nullptr, nullptr);
  return launch_check("upsample_concat_kernel");
}

extern "C" int creste_upsample_concat_split(const float* skip, int Cs, const float* x, int N, int Hi, int Wi, int Cx,
                                            int Ho, int Wo, float rh, float rw, int x_first, const float* amax_a,
                                            const float* amax_b, void* hi, void* lo, float* scal, void* stream) {
  CRESTE_CHECK_ARG(x && hi && scal && amax_a, "creste_upsample_concat_split: null pointer");
  CRESTE_CHECK_ARG((Cs == 0 || skip) && Cs % 4 == 0 && Cx % 4 == 0 && Cx > 0 && (Cs + Cx) % 8 == 0,
                   "creste_upsample_concat_split: channel counts must be multiples of 4 (8 in total)");
  const long long total = (long long)N * Ho * Wo * 32;
  upsample_concat_kernel<true><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(
      skip, Cs, x, N, Hi, Wi, Cx, Ho, Wo, rh, rw, x_first, nullptr, (const unsigned*)amax_a, (const unsigned*)amax_b,
      (uint2*)hi, (uint2*)lo, scal);
  return launch_check("upsample_concat_kernel<split>");
}

extern "C" int creste_maxpool2_concat(const float* const* srcs, const int* chans, int nsrc, int N,
                                      int H, int W, int rows_out, float* out_nhwc, float* out_nchw,
                                      void* stream) {
  CRESTE_CHECK_ARG(srcs && chans && nsrc >= 1 && nsrc <= 3, "creste_maxpool2_concat: 1..3 sources");
  CRESTE_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && rows_out > 0 && rows_out <= H / 2,
                   "creste_maxpool2_concat: bad shape");
  PoolSrcs s;
  int Ct = 0;
  s.n = nsrc;
  for (int i = 0; i < 3; ++i) {
    s.p[i] = i < nsrc ? srcs[i] : nullptr;
    s.c[i] = i < nsrc ? chans[i] : 0;
    Ct += s.c[i];
  }
  const long long total = (long long)N * rows_out * (W / 2) * Ct;
  maxpool2_concat_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(s, N, H, W, rows_out, Ct,
                                                                           out_nhwc, out_nchw);
  return launch_check("maxpool2_concat_kernel");
}

extern "C" int creste_expert_visitation(const void* traj, int is_f64, int B, int T, double map_ds,
                                        int max_steps, int H, int W, float* counts, void* stream) {
  CRESTE_CHECK_ARG(traj && counts && B > 0 && T > 0 && H > 0 && W > 0 && max_steps >= 0,
                   "creste_expert_visitation: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * H * W * sizeof(float), st));
  const long long total = (long long)B * ((long long)(T - 1) * max_steps + 1);
  if (is_f64)
    expert_visitation_kernel<double><<<grid_for(total), 256, 0, st>>>((const double*)traj, B, T, map_ds,
                                                                      max_steps, H, W, counts);
  else
    expert_visitation_kernel<float><<<grid_for(total), 256, 0, st>>>((const float*)traj, B, T,
                                                                     (float)map_ds, max_steps, H, W,
                                                                     counts);
  return launch_check("expert_visitation_kernel");
}
