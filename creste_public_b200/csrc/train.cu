// train.cu -- kernels of the stage-3 (MaxEnt / counterfactual IRL) training step.
//
// The reference trains the reward FCN (creste/models/blocks/conv.py:88-161) with PyTorch autograd,
// including the double backward of the SMODICE gradient penalty
// (creste/utils/loss_utils.py:1208-1217).  Here every primitive of that graph -- and of the graph
// of its backward -- is a CUDA kernel; the host side (creste_public_b200/autograd.py) wraps them in
// torch.autograd.Functions whose backward passes are written in terms of the same set, so the set
// is closed under differentiation:
//
//   conv2d (conv_simt.cu / conv_tc.cu)  <->  conv2d with flipped-transposed weights (dgrad)
//                                       <->  conv2d_wgrad (below)
//   chan_affine (+ReLU)  <->  relu_bwd, chan_dot          (BatchNorm batch statistics are
//   chan_dot             <->  chan_affine                  composed from these two)
//   maxpool2 (layout.cu) <->  maxpool2_bwd <-> maxpool2_gather
//   bilinear upsample (layout.cu) <-> upsample_adjoint
//   row_dot <-> row_scale ; grad_penalty fwd/bwd ; adam_step
//
// All activations NHWC fp32.  Reductions are two-stage with a fixed order (no atomics).
#include "common.cuh"

namespace creste {

static inline int grid_cap(long long total, int threads, int cap) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

// ------------------------------------------------------------------------- chan_affine / relu
// y[pix,c] = act(x[pix,c] * a[c] + b[c]);  a == NULL -> 1, b == NULL -> 0
template <int V>
__global__ void __launch_bounds__(256) chan_affine_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ a,
                                                          const float* __restrict__ b, int C,
                                                          long long n, int relu, float* __restrict__ y,
                                                          unsigned* __restrict__ amax_bits) {
  const long long nv = n / V;
  const int CV = C / V;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * V;
    if (V == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      const float4 aa = a ? __ldg(reinterpret_cast<const float4*>(a + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 bb = b ? __ldg(reinterpret_cast<const float4*>(b + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = fmaf(v.x, aa.x, bb.x); v.y = fmaf(v.y, aa.y, bb.y);
      v.z = fmaf(v.z, aa.z, bb.z); v.w = fmaf(v.w, aa.w, bb.w);
      if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      reinterpret_cast<float4*>(y)[i] = v;
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    } else {
      float v = fmaf(__ldg(x + i), a ? __ldg(a + c) : 1.0f, b ? __ldg(b + c) : 0.0f);
      if (relu) v = fmaxf(v, 0.f);
      y[i] = v;
      m = fmaxf(m, fabsf(v));
    }
  }
  // max|y| for the 3xFP16 operand scale of the conv that consumes y (same reduction as f16_amax_kernel: same bits)
  if (amax_bits) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  }
}

// out = g * (y > 0)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ g,
                                                       const float* __restrict__ y, long long n,
                                                       float* __restrict__ out, unsigned* __restrict__ amax_bits) {
  const long long n4 = n / 4;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + i);
    gv.x = yv.x > 0.f ? gv.x : 0.f; gv.y = yv.y > 0.f ? gv.y : 0.f;
    gv.z = yv.z > 0.f ? gv.z : 0.f; gv.w = yv.w > 0.f ? gv.w : 0.f;
    reinterpret_cast<float4*>(out)[i] = gv;
    m = fmaxf(fmaxf(m, fmaxf(fabsf(gv.x), fabsf(gv.y))), fmaxf(fabsf(gv.z), fabsf(gv.w)));
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float v = __ldg(y + i) > 0.f ? __ldg(g + i) : 0.f;
    out[i] = v;
    m = fmaxf(m, fabsf(v));
  }
  if (amax_bits) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  }
}

// ------------------------------------------------------------------------------------ chan_dot
// part[blk][c] = sum over the block's pixel range of x[pix,c] * (y ? y[pix,c] : 1)
// thread -> (pixel lane, channel group); per-thread sequential sum, then lanes in order.
// Accumulation is in DOUBLE, as PyTorch's CPU batch-norm kernels do (acc_type<float> = double):
// BatchNorm's backward subtracts channel means of the gradient, so the weight gradients upstream
// are differences of nearly cancelling sums and fp32 partial sums cost 3 digits there (measured).
template <int V>
__global__ void __launch_bounds__(256) chan_dot_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ y, int C,
                                                       long long npix, double* __restrict__ part) {
  __shared__ double s_part[256 * V];
  const int CV = C / V;                       // channel groups (<= 256 by the host check)
  const int lanes = 256 / CV;
  const int cg = threadIdx.x % CV, pl = threadIdx.x / CV;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  double acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.0;
  if (pl < lanes) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      if (V == 4) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
        if (y) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + p * C) + cg);
          acc[0] += (double)xv.x * (double)yv.x; acc[1] += (double)xv.y * (double)yv.y;
          acc[2] += (double)xv.z * (double)yv.z; acc[3] += (double)xv.w * (double)yv.w;
        } else {
          acc[0] += (double)xv.x; acc[1] += (double)xv.y; acc[2] += (double)xv.z; acc[3] += (double)xv.w;
        }
      } else {
        const double xv = (double)__ldg(x + p * C + cg);
        acc[0] += y ? xv * (double)__ldg(y + p * C + cg) : xv;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) s_part[threadIdx.x * V + j] = acc[j];
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      double tot = s_part[threadIdx.x * V + j];
      for (int l = 1; l < lanes; ++l) tot += s_part[(l * CV + threadIdx.x) * V + j];
      part[(size_t)blockIdx.x * C + cg * V + j] = tot;
    }
  }
}

// BatchNorm batch statistics in ONE pass: part[blk][0][c] = sum x, part[blk][1][c] = sum x^2 (double)
template <int V>
__global__ void __launch_bounds__(256) chan_stats_kernel(const float* __restrict__ x, int C, long long npix,
                                                         double* __restrict__ part) {
  __shared__ double s_part[2 * 256 * V];
  const int CV = C / V;
  const int lanes = 256 / CV;
  const int cg = threadIdx.x % CV, pl = threadIdx.x / CV;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  double a1[V], a2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) { a1[j] = 0.0; a2[j] = 0.0; }
  if (pl < lanes) {
    for (long long p = p0 + pl; p < p1; p += lanes) {
      float v[V];
      if (V == 4) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + p * C) + cg);
        v[0] = xv.x; v[1 % V] = xv.y; v[2 % V] = xv.z; v[3 % V] = xv.w;
      } else {
        v[0] = __ldg(x + p * C + cg);
      }
#pragma unroll
      for (int j = 0; j < V; ++j) { const double d = (double)v[j]; a1[j] += d; a2[j] += d * d; }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) { s_part[threadIdx.x * V + j] = a1[j]; s_part[256 * V + threadIdx.x * V + j] = a2[j]; }
  __syncthreads();
  if (pl == 0) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      double t1 = s_part[threadIdx.x * V + j], t2 = s_part[256 * V + threadIdx.x * V + j];
      for (int l = 1; l < lanes; ++l) {
        t1 += s_part[(l * CV + threadIdx.x) * V + j];
        t2 += s_part[256 * V + (l * CV + threadIdx.x) * V + j];
      }
      part[((size_t)blockIdx.x * 2) * C + cg * V + j] = t1;
      part[((size_t)blockIdx.x * 2 + 1) * C + cg * V + j] = t2;
    }
  }
}

// second stage of the two-stage reductions, 32 outputs per block: 8 row lanes per output add rows
// lane, lane + 8, ... in order, then the lane sums are added in lane order (fixed order => reproducible).
// A single thread walking all <= 592 rows was 15 % of the IRL step in the round-1 launch list.
template <class Tin, class Tout>
__global__ void __launch_bounds__(256) reduce_rows8_kernel(const Tin* __restrict__ part, int rows, int n, double scale,
                                                           Tout* __restrict__ out) {
  __shared__ double s_sum[8][32];
  const int o = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + o;
  double t = 0.0;
  if (i < n)
    for (int j = rl; j < rows; j += 8) t += (double)part[(size_t)j * n + i];
  s_sum[rl][o] = t;
  __syncthreads();
  if (rl == 0 && i < n) {
    double tot = s_sum[0][o];
#pragma unroll
    for (int l = 1; l < 8; ++l) tot += s_sum[l][o];
    out[i] = (Tout)(tot * scale);
  }
}

__global__ void __launch_bounds__(256) reduce_rows_f64_to_f64_kernel(const double* __restrict__ part, int rows,
                                                                     int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double tot = 0.0;
  for (int j = 0; j < rows; ++j) tot += part[(size_t)j * n + i];
  out[i] = tot;
}

__global__ void __launch_bounds__(256) reduce_rows_f64_kernel(const double* __restrict__ part, int rows,
                                                              int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double tot = 0.0;
  for (int j = 0; j < rows; ++j) tot += part[(size_t)j * n + i];
  out[i] = (float)tot;
}

// out[i] = sum_j part[j][i] in order (double accumulator: the second stage is tiny)
__global__ void __launch_bounds__(256) reduce_rows_kernel(const float* __restrict__ part, int rows,
                                                          int n, float scale, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double tot = 0.0;
  for (int j = 0; j < rows; ++j) tot += (double)__ldg(part + (size_t)j * n + i);
  out[i] = (float)tot * scale;
}

// ------------------------------------------------------------------------------------ maxpool2
// first maximum in (0,0),(0,1),(1,0),(1,1) order, strict '>' (PyTorch max_pool2d tie rule)
__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
  int k = 0; float m = a;
  if (b > m || isnan(b)) { m = b; k = 1; }
  if (c > m || isnan(c)) { m = c; k = 2; }
  if (d > m || isnan(d)) { m = d; k = 3; }
  return k;
}

// mode 0: dx[argmax] = g (others 0)   (x [N,H,W,C], g [N,H/2,W/2,C], out = dx [N,H,W,C])
// mode 1: out[pooled] = gg[argmax]    (gg [N,H,W,C], out [N,H/2,W/2,C])
__global__ void __launch_bounds__(256) maxpool2_route_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ g, int N, int H,
                                                             int W, int C, int mode,
                                                             float* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int n = (int)(pix / ((long long)Wo * Ho));
    const size_t base = (((size_t)n * H + 2 * oy) * W + 2 * ox) * C + c;
    const size_t o[4] = {base, base + C, base + (size_t)W * C, base + (size_t)W * C + C};
    const int k = argmax4(__ldg(x + o[0]), __ldg(x + o[1]), __ldg(x + o[2]), __ldg(x + o[3]));
    if (mode == 0) {
      const float gv = __ldg(g + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) out[o[j]] = (j == k) ? gv : 0.f;
    } else {
      out[i] = __ldg(g + o[k]);
    }
  }
}

// ---------------------------------------------------------------------------- upsample adjoint
// dx[n,iy,ix,c] = sum over output pixels (oy,ox) of wy(oy->iy) * wx(ox->ix) * g[n,oy,ox,c],
// with the weights of creste_upsample_concat (PyTorch bilinear, align_corners=False).
__device__ __forceinline__ float up_weight(int o, float ratio, int isz, int i) {
  const float s = fmaxf(__fsub_rn(__fmul_rn(ratio, __fadd_rn((float)o, 0.5f)), 0.5f), 0.0f);
  const int i0 = (int)s;
  const int i1 = i0 + (i0 < isz - 1 ? 1 : 0);
  const float l1 = __fsub_rn(s, (float)i0), l0 = __fsub_rn(1.0f, l1);
  float w = 0.f;
  if (i0 == i) w += l0;
  if (i1 == i) w += l1;
  return w;
}

// V = 4: one thread = one input pixel x 4 channels (the tap weights depend on the pixel only, so they are
// computed once per 4 channels and every access is 128 bits); per-channel arithmetic order is that of V = 1.
template <int V>
__global__ void __launch_bounds__(256) upsample_adjoint_kernel(const float* __restrict__ g, int N, int Hi,
                                                               int Wi, int C, int Ho, int Wo, float rh,
                                                               float rw, float* __restrict__ dx, int hoist, int ldg) {
  const int CV = C / V;
  const long long total = (long long)N * Hi * Wi * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * V;
    const long long pix = i / CV;
    const int ix = (int)(pix % Wi);
    const int iy = (int)((pix / Wi) % Hi);
    const int n = (int)(pix / ((long long)Wi * Hi));
    int oy0 = (int)floorf(((float)iy - 1.0f + 0.5f) / rh - 0.5f) - 1;
    int oy1 = (int)ceilf(((float)iy + 1.0f + 0.5f) / rh - 0.5f) + 1;
    int ox0 = (int)floorf(((float)ix - 1.0f + 0.5f) / rw - 0.5f) - 1;
    int ox1 = (int)ceilf(((float)ix + 1.0f + 0.5f) / rw - 0.5f) + 1;
    // the last input row / column also receives every clamped (i1 == isz-1) output
    oy0 = max(oy0, 0); ox0 = max(ox0, 0);
    oy1 = (iy == Hi - 1) ? Ho - 1 : min(oy1, Ho - 1);
    ox1 = (ix == Wi - 1) ? Wo - 1 : min(ox1, Wo - 1);
    float acc[V];
#pragma unroll
    for (int u = 0; u < V; ++u) acc[u] = 0.f;
    // the column weights depend on ox only: computed once per thread instead of once per (oy, ox) candidate (a x4
    // adjoint has ~100 candidates per input pixel, 64 of them with a non-zero weight); same values, same order
    constexpr int WXMAX = 16;
    float wxs[WXMAX];
    const int nx = ox1 - ox0 + 1;
    const bool hoisted = hoist && nx <= WXMAX;
    if (hoisted) {
#pragma unroll
      for (int j = 0; j < WXMAX; ++j) wxs[j] = j < nx ? up_weight(ox0 + j, rw, Wi, ix) : 0.f;
    }
    for (int oy = oy0; oy <= oy1; ++oy) {
      const float wy = up_weight(oy, rh, Hi, iy);
      if (wy == 0.f) continue;
      float row[V];
#pragma unroll
      for (int u = 0; u < V; ++u) row[u] = 0.f;
      const float* srow = g + (((size_t)n * Ho + oy) * Wo + ox0) * ldg + c;      // ldg: pixel pitch of g (>= C)
      if (hoisted) {
#pragma unroll
        for (int j = 0; j < WXMAX; ++j) {
          const float wx = wxs[j];
          if (wx == 0.f) continue;
          const float* src = srow + (size_t)j * ldg;
          if (V == 4) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(src));
            row[0] = fmaf(wx, gv.x, row[0]); row[1 % V] = fmaf(wx, gv.y, row[1 % V]);
            row[2 % V] = fmaf(wx, gv.z, row[2 % V]); row[3 % V] = fmaf(wx, gv.w, row[3 % V]);
          } else {
            row[0] = fmaf(wx, __ldg(src), row[0]);
          }
        }
      } else {
        for (int ox = ox0; ox <= ox1; ++ox) {
          const float wx = up_weight(ox, rw, Wi, ix);
          if (wx == 0.f) continue;
          const float* src = srow + (size_t)(ox - ox0) * ldg;
          if (V == 4) {
            const float4 gv = __ldg(reinterpret_cast<const float4*>(src));
            row[0] = fmaf(wx, gv.x, row[0]); row[1 % V] = fmaf(wx, gv.y, row[1 % V]);
            row[2 % V] = fmaf(wx, gv.z, row[2 % V]); row[3 % V] = fmaf(wx, gv.w, row[3 % V]);
          } else {
            row[0] = fmaf(wx, __ldg(src), row[0]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < V; ++u) acc[u] = fmaf(wy, row[u], acc[u]);
    }
    if (V == 4) reinterpret_cast<float4*>(dx)[i] = make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]);
    else dx[i] = acc[0];
  }
}

// --------------------------------------------------------------------------------------- wgrad
// dw[(r*S+s)][c][k] = sum_{n,p,q} x[n, p+r-pad_t, q+s-pad_l, c] * g[n,p,q,k]      (stride 1)
// grid (chunks, R*S): one CTA = one filter tap x one contiguous range of output pixels.  64-pixel
// slabs of x (shifted by the tap, zero outside the image) and g are staged in shared memory; a
// thread owns an 8(c) x 8(k) register tile (64 FMAs per 4 LDS.128 -- FMA-bound, not LDS-bound
// like a 4x4 tile) and one of `lanes` pixel lanes.  Every (chunk, lane) writes its own partial
// [C x K] tile; reduce_rows_kernel adds them in a fixed order.   C, K multiples of 8, <= 64.
constexpr int WG_PIX = 64;
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ x,
                                                    const float* __restrict__ g, int N, int H, int W,
                                                    int C, int K, int R, int S, int pad_t, int pad_l,
                                                    int P, int Q, int lanes, float* __restrict__ part) {
  __shared__ __align__(16) float s_x[WG_PIX][64];
  __shared__ __align__(16) float s_g[WG_PIX][64];
  const int tap = blockIdx.y;
  const int r = tap / S, s = tap - r * S;
  const int tcn = C >> 3, tkn = K >> 3;
  const int per = tcn * tkn;
  const int t = threadIdx.x % per;
  const int tc = t % tcn, tk = t / tcn;
  const int pl = threadIdx.x / per;
  const int c4n = C >> 2, k4n = K >> 2;
  const long long npix = (long long)N * P * Q;
  const long long chunk = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * chunk;
  const long long p1 = p0 + chunk < npix ? p0 + chunk : npix;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (long long base = p0; base < p1; base += WG_PIX) {
    const int cnt = (int)((p1 - base) < WG_PIX ? (p1 - base) : WG_PIX);
    for (int i = threadIdx.x; i < WG_PIX * c4n; i += 256) {
      const int pp = i / c4n, c4 = i - pp * c4n;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pp < cnt) {
        const long long pix = base + pp;
        const int q = (int)(pix % Q);
        const int p = (int)((pix / Q) % P);
        const int n = (int)(pix / ((long long)Q * P));
        const int iy = p + r - pad_t, ix = q + s - pad_l;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * C) + c4);
      }
      *reinterpret_cast<float4*>(&s_x[pp][c4 * 4]) = v;
    }
    for (int i = threadIdx.x; i < WG_PIX * k4n; i += 256) {
      const int pp = i / k4n, k4 = i - pp * k4n;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pp < cnt) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)(base + pp) * K) + k4);
      *reinterpret_cast<float4*>(&s_g[pp][k4 * 4]) = v;
    }
    __syncthreads();
    if (pl < lanes) {
      for (int pp = pl; pp < WG_PIX; pp += lanes) {
        const float4 x0 = *reinterpret_cast<const float4*>(&s_x[pp][tc * 8]);
        const float4 x1 = *reinterpret_cast<const float4*>(&s_x[pp][tc * 8 + 4]);
        const float4 g0 = *reinterpret_cast<const float4*>(&s_g[pp][tk * 8]);
        const float4 g1 = *reinterpret_cast<const float4*>(&s_g[pp][tk * 8 + 4]);
        const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xa[i], ga[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  // reduce the pixel lanes through shared memory (8 accumulators per round, fixed lane order) and
  // write ONE partial [C x K] tile per CTA: part[chunk][tap][C][K]
  float* red = &s_x[0][0];                       // 4096 floats >= lanes * per * 8
  float* dst = part + (((size_t)blockIdx.x * gridDim.y + tap) * C) * K;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __syncthreads();
    if (pl < lanes) {
      float* w = red + ((size_t)pl * per + t) * 8;
      *reinterpret_cast<float4*>(w) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(w + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();
    if (pl == 0) {
      float tot[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tot[j] = red[(size_t)t * 8 + j];
      for (int l = 1; l < lanes; ++l)
#pragma unroll
        for (int j = 0; j < 8; ++j) tot[j] += red[((size_t)l * per + t) * 8 + j];
      float* row = dst + (size_t)(tc * 8 + i) * K + tk * 8;
      *reinterpret_cast<float4*>(row) = make_float4(tot[0], tot[1], tot[2], tot[3]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(tot[4], tot[5], tot[6], tot[7]);
    }
  }
}

// ---- wgrad for R*S > 1: all taps per CTA.  The per-tap kernel above re-reads x and g from L2 once
// per tap (25x for the 5x5 layer).  Here a CTA owns a K slice of 16 output channels and walks
// 8x8-pixel slabs: the x halo tile ((8+R-1) x (8+S-1) pixels x C) and the g tile (64 x 16) are
// staged once per slab and EVERY tap is accumulated from them.  Thread = (tap, 8-channel group of
// x) with an [8 c][16 k] register tile: 128 FMAs per 2 LDS.128 of x + 4 broadcast LDS.128 of g.
// part[chunk][tap][C][K] as above (each CTA fills its 16 columns).
constexpr int WGH_T = 8;           // slab edge (output pixels)
constexpr int WGH_KS = 16;         // output channels per CTA
__global__ void __launch_bounds__(256) wgrad_halo_kernel(const float* __restrict__ x,
                                                         const float* __restrict__ g, int N, int H, int W,
                                                         int C, int K, int R, int S, int pad_t, int pad_l,
                                                         int P, int Q, float* __restrict__ part) {
  extern __shared__ __align__(16) float wsm[];
  const int HT = WGH_T + R - 1, WT = WGH_T + S - 1;     // halo tile
  float* s_x = wsm;                                      // [HT*WT][C]
  float* s_g = wsm + (size_t)HT * WT * C;                // [64][16]
  const int c8n = C >> 3;
  const int nthr = R * S * c8n;                          // active threads
  const int tid = threadIdx.x;
  const bool on = tid < nthr;
  const int tap = on ? tid / c8n : 0, c8 = on ? tid % c8n : 0;
  const int r = tap / S, sft = tap - r * S;
  const int k0 = blockIdx.y * WGH_KS;
  const int tiles_x = (Q + WGH_T - 1) / WGH_T, tiles_y = (P + WGH_T - 1) / WGH_T;
  const int slabs = N * tiles_y * tiles_x;
  const int per = (slabs + gridDim.x - 1) / gridDim.x;
  const int sl0 = blockIdx.x * per, sl1 = min(sl0 + per, slabs);
  const int c4n = C >> 2;
  float acc[8][WGH_KS];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < WGH_KS; ++j) acc[i][j] = 0.f;

  for (int sl = sl0; sl < sl1; ++sl) {
    const int tx = sl % tiles_x, ty = (sl / tiles_x) % tiles_y, n = sl / (tiles_x * tiles_y);
    const int p0 = ty * WGH_T, q0 = tx * WGH_T;
    __syncthreads();
    for (int i = tid; i < HT * WT * c4n; i += blockDim.x) {
      const int pix = i / c4n, c4 = i - pix * c4n;
      const int hy = pix / WT, wx = pix - hy * WT;
      const int iy = p0 + hy - pad_t, ix = q0 + wx - pad_l;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * C) + c4);
      *reinterpret_cast<float4*>(s_x + (size_t)pix * C + c4 * 4) = v;
    }
    for (int i = tid; i < WGH_T * WGH_T * (WGH_KS / 4); i += blockDim.x) {
      const int pix = i / (WGH_KS / 4), k4 = i - pix * (WGH_KS / 4);
      const int py = pix / WGH_T, px = pix - py * WGH_T;
      const int p = p0 + py, q = q0 + px;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < P && q < Q && k0 + k4 * 4 < K)
        v = __ldg(reinterpret_cast<const float4*>(g + (((size_t)n * P + p) * Q + q) * K + k0) + k4);
      *reinterpret_cast<float4*>(s_g + pix * WGH_KS + k4 * 4) = v;
    }
    __syncthreads();
    if (on) {
#pragma unroll 2
      for (int pix = 0; pix < WGH_T * WGH_T; ++pix) {
        const int py = pix >> 3, px = pix & 7;
        const float* xp = s_x + ((size_t)(py + r) * WT + px + sft) * C + c8 * 8;
        const float4 x0 = *reinterpret_cast<const float4*>(xp);
        const float4 x1 = *reinterpret_cast<const float4*>(xp + 4);
        const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float ga[WGH_KS];
#pragma unroll
        for (int j = 0; j < WGH_KS / 4; ++j) {
          const float4 gv = *reinterpret_cast<const float4*>(s_g + pix * WGH_KS + j * 4);
          ga[j * 4] = gv.x; ga[j * 4 + 1] = gv.y; ga[j * 4 + 2] = gv.z; ga[j * 4 + 3] = gv.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < WGH_KS; ++j) acc[i][j] = fmaf(xa[i], ga[j], acc[i][j]);
      }
    }
  }
  if (on) {
    float* dst = part + (((size_t)blockIdx.x * (R * S) + tap) * C + c8 * 8) * K + k0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < WGH_KS; j += 4)
        if (k0 + j < K)
          *reinterpret_cast<float4*>(dst + (size_t)i * K + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
  }
}

// ---------------------------------------------------------------------------- row_dot / row_scale
// out[b] = sum_i x[b,i] * y[b,i] * (m ? m[b,i] : 1)      one CTA per row, fixed order
__global__ void __launch_bounds__(1024) row_dot_kernel(const float* __restrict__ x,
                                                       const float* __restrict__ y,
                                                       const uint8_t* __restrict__ m, long long n,
                                                       float* __restrict__ out) {
  __shared__ float s_red[32];
  const size_t off = (size_t)blockIdx.x * n;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    float v = __ldg(x + off + i) * (y ? __ldg(y + off + i) : 1.0f);
    if (m && !m[off + i]) v = 0.f;
    acc += v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
  }
}

// out[b,i] = x[b,i] * s[b] * (m ? m[b,i] : 1)
__global__ void __launch_bounds__(256) row_scale_kernel(const float* __restrict__ x,
                                                        const float* __restrict__ s,
                                                        const uint8_t* __restrict__ m, long long n,
                                                        long long total, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float v = __ldg(x + i) * __ldg(s + i / n);
    if (m && !m[i]) v = 0.f;
    out[i] = v;
  }
}

// out[b,i] = x[b,i] * (m ? m[b,i] : 1) / (sum_i x[b,i]*m[b,i] + eps)   (loss_utils.py:1142-1146)
__global__ void __launch_bounds__(1024) row_normalize_kernel(const float* __restrict__ x,
                                                             const uint8_t* __restrict__ m, long long n,
                                                             float eps, float* __restrict__ out) {
  __shared__ float s_red[32];
  __shared__ float s_tot;
  const size_t off = (size_t)blockIdx.x * n;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    acc += (m && !m[off + i]) ? 0.f : __ldg(x + off + i);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) s_tot = v + eps;
  }
  __syncthreads();
  const float d = s_tot;
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    out[off + i] = (m && !m[off + i]) ? 0.f : __fdiv_rn(__ldg(x + off + i), d);
}

// ------------------------------------------------------------------------------- grad penalty
// G NCHW [B,C,HW]: per pixel nrm = ||G[b,:,p]||_2 ; fwd: part[blk] = sum (nrm-1)^2 ;
// bwd: dG[b,c,p] = gs * 2 (nrm-1)/nrm * G[b,c,p]  (0 where nrm == 0), gs = g_scalar / (B*HW)
__global__ void __launch_bounds__(256) grad_penalty_kernel(const float* __restrict__ G, int B, int C,
                                                           long long HW, const float* __restrict__ gs,
                                                           float inv_count, float* __restrict__ part,
                                                           float* __restrict__ dG) {
  __shared__ float s_red[8];
  const long long total = (long long)B * HW;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* src = G + (size_t)b * C * HW + p;
    float ss = 0.f;
    for (int c = 0; c < C; ++c) { const float v = __ldg(src + (size_t)c * HW); ss = fmaf(v, v, ss); }
    const float nrm = sqrtf(ss);
    acc += (nrm - 1.0f) * (nrm - 1.0f);
    if (dG) {
      const float k = nrm > 0.f ? __ldg(gs) * inv_count * 2.0f * (nrm - 1.0f) / nrm : 0.f;
      float* dst = dG + (size_t)b * C * HW + p;
      for (int c = 0; c < C; ++c) dst[(size_t)c * HW] = k * __ldg(src + (size_t)c * HW);
    }
  }
  if (part) {
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_red[w];
      part[blockIdx.x] = t;
    }
  }
}

// --------------------------------------------------------------------------------------- Adam
// torch.optim.Adam (no amsgrad, no weight decay): m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
// p -= step_size * m / (sqrt(v)/sqrt(bc2) + eps), step_size = lr / bc1
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   long long n, float b1, float b2, float eps,
                                                   float step_size, float inv_sqrt_bc2, float gscale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);            // lerp form used by torch
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

}  // namespace creste

using namespace creste;

extern "C" int creste_chan_affine(const float* x, const float* a, const float* b, long long npix, int C,
                                  int relu, float* y, void* stream) {
  CRESTE_CHECK_ARG(x && y && npix > 0 && C > 0, "creste_chan_affine: bad args");
  const long long n = npix * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 4 == 0)
    chan_affine_kernel<4><<<grid_cap(n / 4, 256, 148 * 16), 256, 0, st>>>(x, a, b, C, n, relu, y, nullptr);
  else
    chan_affine_kernel<1><<<grid_cap(n, 256, 148 * 16), 256, 0, st>>>(x, a, b, C, n, relu, y, nullptr);
  return launch_check("chan_affine_kernel");
}

extern "C" int creste_chan_affine_amax(const float* x, const float* a, const float* b, long long npix, int C,
                                       int relu, float* y, float* amax_out, void* stream) {
  CRESTE_CHECK_ARG(x && y && amax_out && npix > 0 && C > 0, "creste_chan_affine_amax: bad args");
  const long long n = npix * C;
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, 4, st));
  if (C % 4 == 0)
    chan_affine_kernel<4><<<grid_cap(n / 4, 256, 148 * 16), 256, 0, st>>>(x, a, b, C, n, relu, y, (unsigned*)amax_out);
  else
    chan_affine_kernel<1><<<grid_cap(n, 256, 148 * 16), 256, 0, st>>>(x, a, b, C, n, relu, y, (unsigned*)amax_out);
  return launch_check("chan_affine_kernel");
}

extern "C" int creste_relu_bwd(const float* g, const float* y, long long n, float* out, void* stream) {
  CRESTE_CHECK_ARG(g && y && out && n > 0, "creste_relu_bwd: bad args");
  relu_bwd_kernel<<<grid_cap(n / 4 + 1, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(g, y, n, out, nullptr);
  return launch_check("relu_bwd_kernel");
}

extern "C" int creste_relu_bwd_amax(const float* g, const float* y, long long n, float* out, float* amax_out,
                                    void* stream) {
  CRESTE_CHECK_ARG(g && y && out && amax_out && n > 0, "creste_relu_bwd_amax: bad args");
  CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, 4, (cudaStream_t)stream));
  relu_bwd_kernel<<<grid_cap(n / 4 + 1, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(g, y, n, out,
                                                                                       (unsigned*)amax_out);
  return launch_check("relu_bwd_kernel");
}

static int chan_dot_blocks(long long npix) {
  long long b = npix / 256;
  if (b < 1) b = 1;
  return (int)(b > 592 ? 592 : b);
}

extern "C" size_t creste_chan_dot_workspace_bytes(long long npix, int C) {
  return (size_t)chan_dot_blocks(npix) * C * sizeof(double);
}

extern "C" int creste_chan_dot(const float* x, const float* y, long long npix, int C, float* out, void* ws,
                               size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(x && out && ws && npix > 0 && C > 0 && C <= 1024, "creste_chan_dot: bad args");
  CRESTE_CHECK_ARG(ws_bytes >= creste_chan_dot_workspace_bytes(npix, C), "creste_chan_dot: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = chan_dot_blocks(npix);
  if (C % 4 == 0 && C / 4 <= 256)
    chan_dot_kernel<4><<<blocks, 256, 0, st>>>(x, y, C, npix, (double*)ws);
  else {
    CRESTE_CHECK_ARG(C <= 256, "creste_chan_dot: C %% 4 != 0 needs C <= 256");
    chan_dot_kernel<1><<<blocks, 256, 0, st>>>(x, y, C, npix, (double*)ws);
  }
  int rc = launch_check("chan_dot_kernel");
  if (rc) return rc;
  reduce_rows8_kernel<double, float><<<ceil_div(C, 32), 256, 0, st>>>((const double*)ws, blocks, C, 1.0, out);
  return launch_check("reduce_rows8_kernel");
}

/* out2 DEVICE double[2*C] = {sum_pix x[pix,c]}, {sum_pix x[pix,c]^2}: BatchNorm batch statistics in one
 * pass (double accumulation).  ws >= 2 * creste_chan_dot_workspace_bytes(npix, C). */
extern "C" int creste_chan_stats(const float* x, long long npix, int C, double* out2, void* ws, size_t ws_bytes,
                                 void* stream) {
  CRESTE_CHECK_ARG(x && out2 && ws && npix > 0 && C > 0 && C <= 1024, "creste_chan_stats: bad args");
  CRESTE_CHECK_ARG(ws_bytes >= 2 * creste_chan_dot_workspace_bytes(npix, C), "creste_chan_stats: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = chan_dot_blocks(npix);
  if (C % 4 == 0 && C / 4 <= 256)
    chan_stats_kernel<4><<<blocks, 256, 0, st>>>(x, C, npix, (double*)ws);
  else {
    CRESTE_CHECK_ARG(C <= 256, "creste_chan_stats: C %% 4 != 0 needs C <= 256");
    chan_stats_kernel<1><<<blocks, 256, 0, st>>>(x, C, npix, (double*)ws);
  }
  int rc = launch_check("chan_stats_kernel");
  if (rc) return rc;
  reduce_rows8_kernel<double, double><<<ceil_div(2 * C, 32), 256, 0, st>>>((const double*)ws, blocks, 2 * C, 1.0, out2);
  return launch_check("reduce_rows8_kernel");
}

extern "C" int creste_maxpool2_bwd(const float* x, const float* g, int N, int H, int W, int C, float* dx,
                                   void* stream) {
  CRESTE_CHECK_ARG(x && g && dx && N > 0 && H >= 2 && W >= 2 && C > 0, "creste_maxpool2_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if ((H & 1) || (W & 1)) CRESTE_CUDA(cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * sizeof(float), st));
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  maxpool2_route_kernel<<<grid_cap(total, 256, 148 * 16), 256, 0, st>>>(x, g, N, H, W, C, 0, dx);
  return launch_check("maxpool2_route_kernel");
}

extern "C" int creste_maxpool2_gather(const float* x, const float* gg, int N, int H, int W, int C,
                                      float* out, void* stream) {
  CRESTE_CHECK_ARG(x && gg && out && N > 0 && H >= 2 && W >= 2 && C > 0, "creste_maxpool2_gather: bad args");
  const long long total = (long long)N * (H / 2) * (W / 2) * C;
  maxpool2_route_kernel<<<grid_cap(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, gg, N, H, W, C,
                                                                                        1, out);
  return launch_check("maxpool2_route_kernel");
}

static int upsample_adjoint_launch(const float* g, int ldg, int N, int Hi, int Wi, int C, int Ho, int Wo, float rh, float rw,
                                   float* dx, cudaStream_t st) {
  const long long total = (long long)N * Hi * Wi * C;
  const int hoist = getenv("CRESTE_ADJ_NOHOIST") ? 0 : 1;
  if (C % 4 == 0 && ldg % 4 == 0 && (((uintptr_t)g | (uintptr_t)dx) & 15u) == 0)
    upsample_adjoint_kernel<4><<<grid_cap(total / 4, 256, 148 * 16), 256, 0, st>>>(g, N, Hi, Wi, C, Ho, Wo, rh, rw, dx,
                                                                                   hoist, ldg);
  else
    upsample_adjoint_kernel<1><<<grid_cap(total, 256, 148 * 16), 256, 0, st>>>(g, N, Hi, Wi, C, Ho, Wo, rh, rw, dx, hoist,
                                                                               ldg);
  return launch_check("upsample_adjoint_kernel");
}

extern "C" int creste_upsample_adjoint(const float* g, int N, int Hi, int Wi, int C, int Ho, int Wo, float rh,
                                       float rw, float* dx, void* stream) {
  CRESTE_CHECK_ARG(g && dx && N > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0 && rh > 0 && rw > 0,
                   "creste_upsample_adjoint: bad args");
  return upsample_adjoint_launch(g, C, N, Hi, Wi, C, Ho, Wo, rh, rw, dx, (cudaStream_t)stream);
}

/* the same reading a channel slice [c0, c0 + C) of a wider tensor g [N,Ho,Wo,Cg] in place (the backward of
 * cat([skip, up(x)]): no chan_slice copy of the up-sampled part) */
extern "C" int creste_upsample_adjoint_slice(const float* g, int Cg, int c0, int N, int Hi, int Wi, int C, int Ho, int Wo,
                                             float rh, float rw, float* dx, void* stream) {
  CRESTE_CHECK_ARG(g && dx && N > 0 && Hi > 0 && Wi > 0 && C > 0 && Ho > 0 && Wo > 0 && rh > 0 && rw > 0 && c0 >= 0 &&
                       c0 + C <= Cg, "creste_upsample_adjoint_slice: bad args");
  return upsample_adjoint_launch(g + c0, Cg, N, Hi, Wi, C, Ho, Wo, rh, rw, dx, (cudaStream_t)stream);
}

static bool wgrad_use_halo(const creste_conv_desc* d) {
  return d->R * d->S > 1 && d->R * d->S * (d->C / 8) <= 256 && !getenv("CRESTE_WGRAD_PER_TAP");
}
static int wgrad_halo_chunks(const creste_conv_desc* d) {
  const int slabs = d->N * ceil_div(d->P, WGH_T) * ceil_div(d->Q, WGH_T);
  int chunks = 592 / ceil_div(d->K, WGH_KS);
  if (chunks > slabs) chunks = slabs;
  return chunks < 1 ? 1 : chunks;
}
static int wgrad_chunks(const creste_conv_desc* d) {
  if (wgrad_use_halo(d)) return wgrad_halo_chunks(d);
  const long long npix = (long long)d->N * d->P * d->Q;
  long long chunks = 592 / (d->R * d->S);
  if (chunks < 1) chunks = 1;
  const long long maxc = (npix + 255) / 256;       // >= 256 pixels per chunk
  if (chunks > maxc) chunks = maxc;
  return (int)(chunks < 1 ? 1 : chunks);
}
static int wgrad_lanes(const creste_conv_desc* d) {
  const int per = (d->C / 8) * (d->K / 8);
  const int lanes = 256 / per;
  return lanes > 64 ? 64 : lanes;                  // one pixel of the 64-pixel slab per lane at most
}

extern "C" size_t creste_conv2d_wgrad_workspace_bytes(const creste_conv_desc* d) {
  if (!d || d->C < 8 || d->K < 8) return 0;
  return (size_t)wgrad_chunks(d) * d->R * d->S * d->C * d->K * sizeof(float);
}

/* dw_packed [R*S*C, K] (the creste_conv2d fp32 weight layout with ldw = K) */
extern "C" int creste_conv2d_wgrad(const creste_conv_desc* d, const float* x, const float* g, float* dw_packed,
                                   void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(d && x && g && dw_packed && ws, "creste_conv2d_wgrad: null pointer");
  CRESTE_CHECK_ARG(d->stride == 1, "creste_conv2d_wgrad: stride 1 only");
  CRESTE_CHECK_ARG(d->C % 8 == 0 && d->K % 8 == 0 && d->C <= 64 && d->K <= 64 && d->C > 0 && d->K > 0,
                   "creste_conv2d_wgrad: C, K must be multiples of 8 and <= 64 (got %d, %d)", d->C, d->K);
  CRESTE_CHECK_ARG(ws_bytes >= creste_conv2d_wgrad_workspace_bytes(d), "creste_conv2d_wgrad: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = wgrad_chunks(d), lanes = wgrad_lanes(d);
  int rc;
  if (wgrad_use_halo(d)) {
    const size_t smem = ((size_t)(WGH_T + d->R - 1) * (WGH_T + d->S - 1) * d->C + WGH_T * WGH_T * WGH_KS) * sizeof(float);
    CRESTE_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = (d->R * d->S * (d->C / 8) + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    dim3 grid(chunks, ceil_div(d->K, WGH_KS));
    wgrad_halo_kernel<<<grid, threads, smem, st>>>(x, g, d->N, d->H, d->W, d->C, d->K, d->R, d->S, d->pad_t,
                                                   d->pad_l, d->P, d->Q, (float*)ws);
    rc = launch_check("wgrad_halo_kernel");
  } else {
    dim3 grid(chunks, d->R * d->S);
    wgrad_kernel<<<grid, 256, 0, st>>>(x, g, d->N, d->H, d->W, d->C, d->K, d->R, d->S, d->pad_t, d->pad_l,
                                       d->P, d->Q, lanes, (float*)ws);
    rc = launch_check("wgrad_kernel");
  }
  if (rc) return rc;
  const int n = d->R * d->S * d->C * d->K;
  reduce_rows8_kernel<float, float><<<ceil_div(n, 32), 256, 0, st>>>((const float*)ws, chunks, n, 1.0, dw_packed);
  return launch_check("reduce_rows_kernel");
}

extern "C" int creste_row_dot(const float* x, const float* y, const uint8_t* mask, int B, long long n,
                              float* out, void* stream) {
  CRESTE_CHECK_ARG(x && out && B > 0 && n > 0, "creste_row_dot: bad args");
  row_dot_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(x, y, mask, n, out);
  return launch_check("row_dot_kernel");
}

extern "C" int creste_row_scale(const float* x, const float* s, const uint8_t* mask, int B, long long n,
                                float* out, void* stream) {
  CRESTE_CHECK_ARG(x && s && out && B > 0 && n > 0, "creste_row_scale: bad args");
  const long long total = (long long)B * n;
  row_scale_kernel<<<grid_cap(total, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(x, s, mask, n, total, out);
  return launch_check("row_scale_kernel");
}

extern "C" int creste_row_normalize(const float* x, const uint8_t* mask, int B, long long n, float eps,
                                    float* out, void* stream) {
  CRESTE_CHECK_ARG(x && out && B > 0 && n > 0, "creste_row_normalize: bad args");
  row_normalize_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(x, mask, n, eps, out);
  return launch_check("row_normalize_kernel");
}

static int gp_blocks(long long total) { return grid_cap(total, 256, 592); }

extern "C" size_t creste_grad_penalty_workspace_bytes(int B, long long HW) {
  return (size_t)gp_blocks((long long)B * HW) * sizeof(float);
}

/* penalty_out (device scalar) = mean over (b,pixel) of (||G[b,:,pixel]||_2 - 1)^2, G NCHW */
extern "C" int creste_grad_penalty(const float* G, int B, int C, long long HW, float* penalty_out, void* ws,
                                   size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(G && penalty_out && ws && B > 0 && C > 0 && HW > 0, "creste_grad_penalty: bad args");
  CRESTE_CHECK_ARG(ws_bytes >= creste_grad_penalty_workspace_bytes(B, HW), "creste_grad_penalty: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = gp_blocks((long long)B * HW);
  grad_penalty_kernel<<<blocks, 256, 0, st>>>(G, B, C, HW, nullptr, 0.f, (float*)ws, nullptr);
  int rc = launch_check("grad_penalty_kernel");
  if (rc) return rc;
  // sum of the block partials in order, then / count
  reduce_rows_kernel<<<1, 256, 0, st>>>((const float*)ws, blocks, 1, 1.0f / (float)((long long)B * HW),
                                        penalty_out);
  return launch_check("reduce_rows_kernel");
}

/* dG = g_scalar * d penalty / dG   (g_scalar: device scalar) */
extern "C" int creste_grad_penalty_bwd(const float* G, const float* g_scalar, int B, int C, long long HW,
                                       float* dG, void* stream) {
  CRESTE_CHECK_ARG(G && g_scalar && dG && B > 0 && C > 0 && HW > 0, "creste_grad_penalty_bwd: bad args");
  const float inv = 1.0f / (float)((long long)B * HW);
  grad_penalty_kernel<<<gp_blocks((long long)B * HW), 256, 0, (cudaStream_t)stream>>>(G, B, C, HW, g_scalar,
                                                                                    inv, nullptr, dG);
  return launch_check("grad_penalty_kernel");
}

extern "C" int creste_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1,
                                float b2, float eps, int step, float grad_scale, void* stream) {
  CRESTE_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "creste_adam_step: bad args");
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  adam_kernel<<<grid_cap(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, b1, b2, eps, (float)((double)lr / bc1), (float)(1.0 / sqrt(bc2)), grad_scale);
  return launch_check("adam_kernel");
}

// ------------------------------------------------------------------ stage-1 loss values (eval)
// CrossEntropyDepth + SmoothL1Depth forward values in ONE pass over the depth logits
// (reference creste/utils/loss_utils.py:477-573, bin_depths creste/utils/depth_utils.py:346-383,
// mode "UD").  One thread per pixel, coalesced across pixels for every class plane of the NCHW
// logits.  acc[0] = sum of -log_softmax(logits)[gt_bin] over valid pixels, acc[1] = #valid,
// acc[2] = #(argmax == gt_bin), acc[3] = sum of smooth_l1(pred_bins - label_m) over valid pixels.
namespace creste {
__global__ void __launch_bounds__(256) stage1_depth_loss_kernel(const float* __restrict__ logits,
                                                                const long long* __restrict__ pred_bins,
                                                                const float* __restrict__ label_mm, int N,
                                                                int D, long long HW, float dmin, float bin_size,
                                                                float beta, double* __restrict__ acc) {
  __shared__ double s_red[4][8];
  double ce = 0.0, nv = 0.0, nc = 0.0, sl = 0.0;
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const float d = __ldg(label_mm + i);
    const float idxf = __fdiv_rn(__fsub_rn(d, dmin), bin_size);
    const bool bad = (idxf < 0.0f) || (idxf > (float)D) || !isfinite(idxf);
    const long long bin = bad ? (long long)D : (long long)idxf;        // truncation, as .type(int64)
    if (bin == D) continue;
    const float* src = logits + (size_t)n * D * HW + p;
    float m = -INFINITY;
    int am = 0;
    for (int k = 0; k < D; ++k) {
      const float v = __ldg(src + (size_t)k * HW);
      if (v > m) { m = v; am = k; }
    }
    float ssum = 0.f;
    for (int k = 0; k < D; ++k) ssum += expf(__ldg(src + (size_t)k * HW) - m);
    ce += (double)(logf(ssum) + m - __ldg(src + (size_t)bin * HW));
    nv += 1.0;
    nc += (am == (int)bin) ? 1.0 : 0.0;
    const float diff = fabsf((float)__ldg(pred_bins + i) - d / 1000.0f);
    sl += (double)(diff < beta ? 0.5f * diff * diff / beta : diff - 0.5f * beta);
  }
  double vals[4] = {ce, nv, nc, sl};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double v = vals[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[j][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, t);
  }
}

// acc[0] = sum (pred - gt)^2 over elements with !isinf(gt), acc[1] = their count (MSELoss, :606-647)
__global__ void __launch_bounds__(256) masked_mse_kernel(const float* __restrict__ pred,
                                                         const float* __restrict__ gt, long long n,
                                                         double* __restrict__ acc) {
  __shared__ double s_red[2][8];
  double ss = 0.0, cnt = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float g = __ldg(gt + i);
    if (isinf(g)) continue;
    const float d = __ldg(pred + i) - g;
    ss += (double)d * (double)d;
    cnt += 1.0;
  }
  double vals[2] = {ss, cnt};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    double v = vals[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[j][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[threadIdx.x][w];
    atomicAdd(acc + threadIdx.x, t);
  }
}
}  // namespace creste

extern "C" int creste_stage1_depth_losses(const float* logits_nchw, const long long* pred_bins,
                                          const float* label_mm, int N, int D, long long HW, float depth_min,
                                          float depth_max, float beta, double* acc4, void* stream) {
  CRESTE_CHECK_ARG(logits_nchw && pred_bins && label_mm && acc4 && N > 0 && D > 0 && HW > 0 && beta > 0,
                   "creste_stage1_depth_losses: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(acc4, 0, 4 * sizeof(double), st));
  const float bin_size = (float)(((double)depth_max - (double)depth_min) / (double)D);
  stage1_depth_loss_kernel<<<grid_cap((long long)N * HW, 256, 148 * 8), 256, 0, st>>>(
      logits_nchw, pred_bins, label_mm, N, D, HW, depth_min, bin_size, beta, acc4);
  return launch_check("stage1_depth_loss_kernel");
}

extern "C" int creste_masked_mse(const float* pred, const float* gt, long long n, double* acc2, void* stream) {
  CRESTE_CHECK_ARG(pred && gt && acc2 && n > 0, "creste_masked_mse: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st));
  masked_mse_kernel<<<grid_cap(n, 256, 148 * 8), 256, 0, st>>>(pred, gt, n, acc2);
  return launch_check("masked_mse_kernel");
}
