// layout.cu -- memory-bound glue kernels (NHWC): layout shuffles at the module boundary,
// bilinear upsample + channel concat, 2x2 max-pool + concat + crop, expert-visitation raster.
// All are pure HBM streaming kernels: float4-vectorised, channel-contiguous, grid-stride.
#include <cuda_fp16.h>
#include <stdlib.h>

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

// NCHW -> NHWC for C == 4 (the RGB-D input): one pixel per thread, four coalesced plane reads, one float4 store
__global__ void __launch_bounds__(256) nchw4_to_nhwc_kernel(const float* __restrict__ in, long long HW, long long total,
                                                            float4* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const float* b = in + n * 4 * HW + p;
    out[i] = make_float4(__ldg(b), __ldg(b + HW), __ldg(b + 2 * HW), __ldg(b + 3 * HW));
  }
}

// NCHW -> NHWC for a handful of channels (C <= 8, C != 4: the 6- and 2-channel head predictions that enter the reward
// net at 1024 x 512): one pixel per thread, C coalesced plane reads, C consecutive floats stored.  The 32 x 32 tile of
// transpose_kernel has C valid rows of 32 there (190 us for 100 MB).
template <int C>
__global__ void __launch_bounds__(256) nchw_small_to_nhwc_kernel(const float* __restrict__ in, long long HW,
                                                                 long long total, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const float* b = in + n * C * HW + p;
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __ldg(b + c * HW);
    float* o = out + i * C;
    if (C % 2 == 0) {
#pragma unroll
      for (int c = 0; c < C; c += 2) *reinterpret_cast<float2*>(o + c) = make_float2(v[c], v[c + 1 < C ? c + 1 : c]);
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = v[c];
    }
  }
}

// 128-bit form of transpose_kernel for rows % 4 == 0 && cols % 4 == 0: a 64 x 64 tile, float4 loads along the
// columns and float4 stores along the rows (the scalar form moved 4 bytes per lane per instruction and ran the
// 126 MB layout changes at the module boundary at ~0.25 of the HBM rate)
__global__ void __launch_bounds__(256) transpose4_kernel(const float* __restrict__ in, int rows, int cols,
                                                         float* __restrict__ out) {
  __shared__ float tile[64][65];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = t + k * 256;               // 64 rows x 16 float4
    const int r = f >> 4, c4 = f & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < rows && c0 + c4 * 4 < cols)
      v = __ldg(reinterpret_cast<const float4*>(in + base + (size_t)(r0 + r) * cols + c0 + c4 * 4));
    tile[r][c4 * 4 + 0] = v.x; tile[r][c4 * 4 + 1] = v.y; tile[r][c4 * 4 + 2] = v.z; tile[r][c4 * 4 + 3] = v.w;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = t + k * 256;               // 64 cols x 16 float4 (along the rows)
    const int c = f >> 4, r4 = f & 15;
    if (c0 + c < cols && r0 + r4 * 4 < rows) {
      const float4 v = make_float4(tile[r4 * 4 + 0][c], tile[r4 * 4 + 1][c], tile[r4 * 4 + 2][c], tile[r4 * 4 + 3][c]);
      *reinterpret_cast<float4*>(out + base + (size_t)(c0 + c) * rows + r0 + r4 * 4) = v;
    }
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
  // weights transposed to [C][KO]: the KO products of one input channel read their weights as float4 broadcasts
  // (one shared-memory load per 4 FFMA; the [KO][C] form issued one load per FFMA and made the kernel LSU-bound)
  extern __shared__ __align__(16) float s_w[];          // [C][KO], then the per-warp tiles [8][2][32][33]
  float (*tile)[2][32][33] = reinterpret_cast<float (*)[2][32][33]>(s_w + KO * C);
  for (int i = threadIdx.x; i < KO * C; i += blockDim.x) {
    const int k = i / C, c = i - k * C;
    s_w[c * KO + k] = (k < K) ? w[i] : 0.0f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // stage one 32-pixel x 32-channel chunk into tile buffer `buf` with cp.async (no registers, no stall): the next
  // chunk is in flight while the current one is multiplied (the synchronous form waited a full DRAM latency per
  // batch of 8 rows with 16 warps per SM: 6.3 of 10 issue slots idle on long-scoreboard stalls, ncu round 2)
  auto stage = [&](long long p0, int c0, int buf) {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      float* dst = &tile[warp][buf][r][lane];
      if (p0 + r < M) {
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(x + (size_t)(p0 + r) * C + c0 + lane) : "memory");
      } else {
        *dst = 0.0f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const long long pstep = (long long)gridDim.x * 256;
  const int nch = C / 32;
  long long p0 = ((long long)blockIdx.x * 8 + warp) * 32;
  int buf = 0;
  if (p0 < M) stage(p0, 0, 0);
  for (; p0 < M; p0 += pstep) {
    const long long pix = p0 + lane;                   // this lane's pixel in the compute / store phases
    const bool ok = pix < M;
    const long long img = ok ? pix / PQ : 0, pp = ok ? pix - img * PQ : 0;
    float acc[KO];
#pragma unroll
    for (int k = 0; k < KO; ++k) acc[k] = 0.0f;
    for (int ch = 0; ch < nch; ++ch) {
      const int c0 = ch * 32;
      // prefetch the next chunk (of this pixel group, or the first chunk of the next one)
      if (ch + 1 < nch) stage(p0, c0 + 32, buf ^ 1);
      else if (p0 + pstep < M) stage(p0 + pstep, 0, buf ^ 1);
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
#pragma unroll 8
      for (int c = 0; c < 32; ++c) {
        const float xv = tile[warp][buf][lane][c];
        const float* wr = s_w + (c0 + c) * KO;
        if constexpr (KO % 4 == 0) {
#pragma unroll
          for (int k = 0; k < KO; k += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
            acc[k] = fmaf(xv, w4.x, acc[k]); acc[k + 1] = fmaf(xv, w4.y, acc[k + 1]);
            acc[k + 2] = fmaf(xv, w4.z, acc[k + 2]); acc[k + 3] = fmaf(xv, w4.w, acc[k + 3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < KO; ++k) acc[k] = fmaf(xv, wr[k], acc[k]);
        }
        if (x_nchw && ok) x_nchw[((size_t)img * C + c0 + c) * PQ + pp] = xv;
      }
      __syncwarp();
      buf ^= 1;
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
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
        uint2 h2, l2;
        split4_f16(v.x * sc, v.y * sc, v.z * sc, v.w * sc, h2, l2);
        hi[pix * c4t + c4] = h2;
        if (lo) lo[pix * c4t + c4] = l2;
      }
    }
  }
}

// ---- the same for an integer factor F in {2, 4} (Ho = F * Hi, Wo = F * Wi, ratio 1 / F): one warp per INPUT pixel.
// With src = (dst + 0.5) / F - 0.5 the F x F output block of input pixel (by, bx) takes its taps from the 3 x 3
// neighbourhood only (first half of the block: (b - 1, b), second half: (b, b + 1)), so a lane loads 9 float4 once
// and produces F * F outputs instead of 4 loads per output, and the horizontal interpolation of an input row is
// shared by the output rows that use it.  Bit-identical to upsample_concat_kernel: the tap weights come from the
// same float formula per output row / column, the two-level interpolation is evaluated as the same
// fma(l0, a, l1 * b) pairs, and at the borders the clamped neighbour carries weight exactly 0 or duplicates the tap
// (1 * a + 0 * b = a for finite b).
template <int F, bool SPLIT>
__global__ void __launch_bounds__(256, 3) upsample_block_kernel(const float* __restrict__ skip, int Cs,
                                                             const float* __restrict__ x, int N, int Hi, int Wi, int Cx,
                                                             int x_first, float* __restrict__ out,
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
  constexpr float R = 1.0f / F;
  const int Ho = Hi * F, Wo = Wi * F;
  const int Ct = Cs + Cx;
  const int c4t = Ct / 4, cs4 = Cs / 4, cx4 = Cx / 4;
  const int lane = threadIdx.x & 31;
  const long long nin = (long long)N * Hi * Wi;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  auto emit = [&](long long pix, int c4, const float4 v) {
    if (!SPLIT) {
      reinterpret_cast<float4*>(out + pix * Ct)[c4] = v;
    } else {
      uint2 h2, l2;
      split4_f16(v.x * sc, v.y * sc, v.z * sc, v.w * sc, h2, l2);
      hi[pix * c4t + c4] = h2;
      if (lo) lo[pix * c4t + c4] = l2;
    }
  };
  for (long long item = warp0; item < nin; item += nwarps) {
    const int bx = (int)(item % Wi);
    const long long t = item / Wi;
    const int by = (int)(t % Hi);
    const int n = (int)(t / Hi);
    const int ry[3] = {max(by - 1, 0), by, min(by + 1, Hi - 1)};
    const int rx[3] = {max(bx - 1, 0), bx, min(bx + 1, Wi - 1)};
    const float* b = x + (size_t)n * Hi * Wi * Cx;
    const long long opix0 = ((long long)n * Ho + (long long)by * F) * Wo + (long long)bx * F;
    for (int c4 = lane; c4 < c4t; c4 += 32) {
      const bool is_skip = x_first ? (c4 >= cx4) : (c4 < cs4);
      if (is_skip) {
        const int sc4 = x_first ? c4 - cx4 : c4;
#pragma unroll 1
        for (int i = 0; i < F; ++i)
#pragma unroll
          for (int k = 0; k < F; ++k) {
            const long long pix = opix0 + (long long)i * Wo + k;
            emit(pix, c4, __ldg(reinterpret_cast<const float4*>(skip + pix * Cs) + sc4));
          }
        continue;
      }
      const int cc = x_first ? c4 : c4 - cs4;
      float4 V[3][3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q)
          V[r][q] = __ldg(reinterpret_cast<const float4*>(b + ((size_t)ry[r] * Wi + rx[q]) * Cx) + cc);
      // the tap weights are recomputed per output column / row (six FP instructions) instead of being kept in
      // 4 * F registers, and only the half index is unrolled: the fully unrolled form needed 255 registers
      constexpr int HALF = F / 2;
#pragma unroll
      for (int qh = 0; qh < 2; ++qh) {
#pragma unroll 1
        for (int kk = 0; kk < HALF; ++kk) {
          const int k = qh * HALF + kk;
          const float sx = fmaxf(__fsub_rn(__fmul_rn(R, __fadd_rn((float)(bx * F + k), 0.5f)), 0.5f), 0.0f);
          const float l1w = __fsub_rn(sx, (float)(int)sx), l0w = __fsub_rn(1.0f, l1w);
          float4 Hr[3];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float4 a = V[r][qh], c = V[r][qh + 1];
            Hr[r].x = __fmaf_rn(l0w, a.x, __fmul_rn(l1w, c.x));
            Hr[r].y = __fmaf_rn(l0w, a.y, __fmul_rn(l1w, c.y));
            Hr[r].z = __fmaf_rn(l0w, a.z, __fmul_rn(l1w, c.z));
            Hr[r].w = __fmaf_rn(l0w, a.w, __fmul_rn(l1w, c.w));
          }
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
#pragma unroll 1
            for (int ii = 0; ii < HALF; ++ii) {
              const int i = rh * HALF + ii;
              const float sy = fmaxf(__fsub_rn(__fmul_rn(R, __fadd_rn((float)(by * F + i), 0.5f)), 0.5f), 0.0f);
              const float l1h = __fsub_rn(sy, (float)(int)sy), l0h = __fsub_rn(1.0f, l1h);
              float4 v;
              v.x = __fmaf_rn(l0h, Hr[rh].x, __fmul_rn(l1h, Hr[rh + 1].x));
              v.y = __fmaf_rn(l0h, Hr[rh].y, __fmul_rn(l1h, Hr[rh + 1].y));
              v.z = __fmaf_rn(l0h, Hr[rh].z, __fmul_rn(l1h, Hr[rh + 1].z));
              v.w = __fmaf_rn(l0h, Hr[rh].w, __fmul_rn(l1h, Hr[rh + 1].w));
              emit(opix0 + (long long)i * Wo + k, c4, v);
            }
          }
        }
      }
    }
  }
}

// integer-factor dispatch shared by the two entry points; returns false if the shape is not an exact x2 / x4
template <bool SPLIT>
static bool upsample_block_launch(const float* skip, int Cs, const float* x, int N, int Hi, int Wi, int Cx, int Ho, int Wo,
                                  float rh, float rw, int x_first, float* out, const unsigned* amax_a,
                                  const unsigned* amax_b, uint2* hi, uint2* lo, float* scal, cudaStream_t st) {
  if (getenv("CRESTE_NO_UPSAMPLE_BLOCK")) return false;
  int F = 0;
  if (Ho == 2 * Hi && Wo == 2 * Wi && rh == 0.5f && rw == 0.5f) F = 2;
  if (Ho == 4 * Hi && Wo == 4 * Wi && rh == 0.25f && rw == 0.25f) F = 4;
  if (!F) return false;
  const long long total = (long long)N * Hi * Wi * 32;      // one warp per input pixel
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (F == 2)
    upsample_block_kernel<2, SPLIT><<<(int)blocks, 256, 0, st>>>(skip, Cs, x, N, Hi, Wi, Cx, x_first, out, amax_a, amax_b,
                                                                 hi, lo, scal);
  else
    upsample_block_kernel<4, SPLIT><<<(int)blocks, 256, 0, st>>>(skip, Cs, x, N, Hi, Wi, Cx, x_first, out, amax_a, amax_b,
                                                                 hi, lo, scal);
  return true;
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
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  if (C == 4 && al16(out)) {
    const long long total = (long long)N * H * W;
    nchw4_to_nhwc_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(in, (long long)H * W, total,
                                                                          reinterpret_cast<float4*>(out));
    return launch_check("nchw4_to_nhwc_kernel");
  }
  if (C <= 8 && C != 4 && C != 8 && (((uintptr_t)out & 7u) == 0 || C % 2) && !getenv("CRESTE_NO_SMALL_TRANSPOSE")) {
    const long long total = (long long)N * H * W, HW = (long long)H * W;
    cudaStream_t st = (cudaStream_t)stream;
    switch (C) {
      case 1: nchw_small_to_nhwc_kernel<1><<<grid_for(total), 256, 0, st>>>(in, HW, total, out); break;
      case 2: nchw_small_to_nhwc_kernel<2><<<grid_for(total), 256, 0, st>>>(in, HW, total, out); break;
      case 3: nchw_small_to_nhwc_kernel<3><<<grid_for(total), 256, 0, st>>>(in, HW, total, out); break;
      case 5: nchw_small_to_nhwc_kernel<5><<<grid_for(total), 256, 0, st>>>(in, HW, total, out); break;
      case 6: nchw_small_to_nhwc_kernel<6><<<grid_for(total), 256, 0, st>>>(in, HW, total, out); break;
      default: nchw_small_to_nhwc_kernel<7><<<grid_for(total), 256, 0, st>>>(in, HW, total, out); break;
    }
    return launch_check("nchw_small_to_nhwc_kernel");
  }
  if (C % 4 == 0 && (H * W) % 4 == 0 && al16(in) && al16(out)) {
    dim3 grid4(ceil_div(H * W, 64), ceil_div(C, 64), N);
    transpose4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(in, C, H * W, out);
    return launch_check("transpose4_kernel");
  }
  dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), N);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, C, H * W, out);
  return launch_check("transpose_kernel");
}

extern "C" int creste_nhwc_to_nchw(const float* in, int N, int H, int W, int C, float* out,
                                   void* stream) {
  CRESTE_CHECK_ARG(in && out && N > 0 && C > 0 && H > 0 && W > 0, "creste_nhwc_to_nchw: bad args");
  if (C % 4 == 0 && (H * W) % 4 == 0 && ((uintptr_t)in & 15u) == 0 && ((uintptr_t)out & 15u) == 0) {
    dim3 grid4(ceil_div(C, 64), ceil_div(H * W, 64), N);
    transpose4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(in, H * W, C, out);
    return launch_check("transpose4_kernel");
  }
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
  do {                                                                                                          \
    const size_t smem = (size_t)KO * C * sizeof(float) + 8 * 2 * 32 * 33 * sizeof(float);                       \
    CRESTE_CUDA(cudaFuncSetAttribute(proj_head_kernel<KO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    proj_head_kernel<KO><<<(int)blocks, 256, smem, st>>>(x, w, bias, M, C, PQ, pred_nhwc, pred_nchw, x_nchw, K); \
  } while (0)
  if (K <= 2) CRESTE_PROJ(2);
  else if (K <= 8) CRESTE_PROJ(8);
  else if (K <= 16) CRESTE_PROJ(16);
  else CRESTE_PROJ(32);
#undef CRESTE_PROJ
  return launch_check("proj_head_kernel");
}

extern "C" int creste_upsample_concat(const float* skip, int Cs, const float* x, int N, int Hi,
                                      int Wi, int Cx, int Ho, int Wo, float rh, float rw, int x_first,
                                      float* out, void* stream) {
  CRESTE_CHECK_ARG(x && out, "creste_upsample_concat: null pointer");
  CRESTE_CHECK_ARG((Cs == 0 || skip) && Cs % 4 == 0 && Cx % 4 == 0 && Cx > 0,
                   "creste_upsample_concat: channel counts must be multiples of 4");
  if (upsample_block_launch<false>(skip, Cs, x, N, Hi, Wi, Cx, Ho, Wo, rh, rw, x_first, out, nullptr, nullptr, nullptr,
                                   nullptr, nullptr, (cudaStream_t)stream))
    return launch_check("upsample_block_kernel");
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
  if (upsample_block_launch<true>(skip, Cs, x, N, Hi, Wi, Cx, Ho, Wo, rh, rw, x_first, nullptr, (const unsigned*)amax_a,
                                  (const unsigned*)amax_b, (uint2*)hi, (uint2*)lo, scal, (cudaStream_t)stream))
    return launch_check("upsample_block_kernel<split>");
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
