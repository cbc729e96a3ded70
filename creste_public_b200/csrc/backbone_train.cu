// backbone_train.cu -- kernels of the stage-1 (distillation) training step of the RGB-D backbone.
//
// The reference trains DistillationBackbone (creste/models/distillation.py:145-207: EfficientNet-B0
// trunk + U-Net decoder, creste/models/blocks/effnet.py:8-98, depth head and dino head,
// creste/models/blocks/conv.py:5-32) with PyTorch autograd against CrossEntropyDepth + MSELoss
// (creste/utils/loss_utils.py:477-527, 606-647) in creste/train_pefree.py:76-106.  The dense convs
// of that graph reuse conv2d / conv2d_wgrad (conv_tc.cu, conv_simt.cu, train.cu); this file holds
// everything else the train-mode graph and its backward need:
//
//   chan_moments               BatchNorm batch statistics, any C % 4 == 0 (double accumulators)
//   chan_affine_act            y = act(x * a[c] + b[c]), act in {none, relu, swish}
//   bn_act_bwd                 gu = g * act'(x*a+b)  +  (sum gu, sum gu*x) per channel, one pass
//   chan_axpby                 dx = gu*p[c] + x*q[c] + r[c]      (BatchNorm backward, one pass)
//   dwconv_fwd / dgrad / wgrad depthwise k x k conv with TF-'SAME' (asymmetric) padding, stride 1/2
//   sample_dot / sample_affine squeeze-excite pooling / gating and their adjoints ([B,C] vectors)
//   act / act_bwd              swish, sigmoid on the tiny SE tensors
//   add_scaled                 identity skip + drop-connect:  out = inp + x * s[b]
//   chan_slice                 channel range copy (adjoint of the decoder's concat)
//   wgrad_strided              weight gradient of the strided C=4 stem conv
//   ce_depth_bwd, masked_mse_bwd   loss gradients
//
// All activations NHWC fp32.  The reductions of this file are two-stage with a fixed order (no atomics).
#include <stdlib.h>

#include "common.cuh"

namespace creste {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SWISH = 2, ACT_SIGMOID = 3 };

static inline int grid_cap2(long long total, int threads, int cap) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

__device__ __forceinline__ float sigmoid_f(float u) { return 1.0f / (1.0f + expf(-u)); }
__device__ __forceinline__ float act_fwd(float u, int act) {
  if (act == ACT_RELU) return fmaxf(u, 0.f);
  if (act == ACT_SWISH) return u * sigmoid_f(u);
  if (act == ACT_SIGMOID) return sigmoid_f(u);
  return u;
}
// d act(u) / du
__device__ __forceinline__ float act_grad(float u, int act) {
  if (act == ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == ACT_SWISH) { const float s = sigmoid_f(u); return s * (1.f + u * (1.f - s)); }
  if (act == ACT_SIGMOID) { const float s = sigmoid_f(u); return s * (1.f - s); }
  return 1.f;
}

// -------------------------------------------------------------------------------- elementwise
__global__ void __launch_bounds__(256) chan_affine_act_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ a,
                                                              const float* __restrict__ b, int C, long long n,
                                                              int act, float* __restrict__ y,
                                                              unsigned* __restrict__ amax_bits) {
  const long long nv = n / 4;
  const int CV = C / 4;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 aa = a ? __ldg(reinterpret_cast<const float4*>(a + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bb = b ? __ldg(reinterpret_cast<const float4*>(b + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = act_fwd(fmaf(v.x, aa.x, bb.x), act); v.y = act_fwd(fmaf(v.y, aa.y, bb.y), act);
    v.z = act_fwd(fmaf(v.z, aa.z, bb.z), act); v.w = act_fwd(fmaf(v.w, aa.w, bb.w), act);
    reinterpret_cast<float4*>(y)[i] = v;
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  // max|y| for the consumer's 3xFP16 operand scale (the same reduction as f16_amax_kernel: same bits)
  if (amax_bits) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  }
}

// out = u*p[c] + x*q[c] + r[c]
__global__ void __launch_bounds__(256) chan_axpby_kernel(const float* __restrict__ u, const float* __restrict__ x,
                                                         const float* __restrict__ p, const float* __restrict__ q,
                                                         const float* __restrict__ r, int C, long long n,
                                                         float* __restrict__ out, unsigned* __restrict__ amax_bits) {
  const long long nv = n / 4;
  const int CV = C / 4;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    const float4 uv = __ldg(reinterpret_cast<const float4*>(u) + i);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 pp = __ldg(reinterpret_cast<const float4*>(p + c));
    const float4 qq = __ldg(reinterpret_cast<const float4*>(q + c));
    const float4 rr = __ldg(reinterpret_cast<const float4*>(r + c));
    float4 o;
    o.x = fmaf(uv.x, pp.x, fmaf(xv.x, qq.x, rr.x)); o.y = fmaf(uv.y, pp.y, fmaf(xv.y, qq.y, rr.y));
    o.z = fmaf(uv.z, pp.z, fmaf(xv.z, qq.z, rr.z)); o.w = fmaf(uv.w, pp.w, fmaf(xv.w, qq.w, rr.w));
    reinterpret_cast<float4*>(out)[i] = o;
    m = fmaxf(fmaxf(m, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
  }
  if (amax_bits) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  }
}

// out = (g * act'(x*a[c] + b[c])) * p[c] + x*q[c] + r[c]: the second half of the BatchNorm(+act) backward with gu
// recomputed from g instead of read back (the first half then does not store it): same expressions as BnActBwdOp
// followed by chan_axpby_kernel, same bits, 20 instead of 24 bytes per element over the two passes.
__global__ void __launch_bounds__(256) chan_axpby_act_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                             const float* __restrict__ a, const float* __restrict__ b,
                                                             int act, const float* __restrict__ p,
                                                             const float* __restrict__ q, const float* __restrict__ r,
                                                             int C, long long n, float* __restrict__ out,
                                                             unsigned* __restrict__ amax_bits) {
  const long long nv = n / 4;
  const int CV = C / 4;
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    float4 gv = __ldg(reinterpret_cast<const float4*>(g) + i);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float4 aa = __ldg(reinterpret_cast<const float4*>(a + c));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
    gv.x *= act_grad(fmaf(xv.x, aa.x, bb.x), act); gv.y *= act_grad(fmaf(xv.y, aa.y, bb.y), act);
    gv.z *= act_grad(fmaf(xv.z, aa.z, bb.z), act); gv.w *= act_grad(fmaf(xv.w, aa.w, bb.w), act);
    const float4 pp = __ldg(reinterpret_cast<const float4*>(p + c));
    const float4 qq = __ldg(reinterpret_cast<const float4*>(q + c));
    const float4 rr = __ldg(reinterpret_cast<const float4*>(r + c));
    float4 o;
    o.x = fmaf(gv.x, pp.x, fmaf(xv.x, qq.x, rr.x)); o.y = fmaf(gv.y, pp.y, fmaf(xv.y, qq.y, rr.y));
    o.z = fmaf(gv.z, pp.z, fmaf(xv.z, qq.z, rr.z)); o.w = fmaf(gv.w, pp.w, fmaf(xv.w, qq.w, rr.w));
    reinterpret_cast<float4*>(out)[i] = o;
    m = fmaxf(fmaxf(m, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
  }
  if (amax_bits) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
  }
}

// out[b,pix,c] = (x ? x[b,pix,c] * (a ? a[b,c] : 1) : 0) + (bb ? bb[b,c] : 0)
__global__ void __launch_bounds__(256) sample_affine_kernel(const float* __restrict__ x,
                                                            const float* __restrict__ a,
                                                            const float* __restrict__ bb, long long HW, int C,
                                                            long long n, float* __restrict__ out) {
  const long long nv = n / 4;
  const int CV = C / 4;
  const long long per = HW * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    const long long s = i / per;
    float4 v = x ? __ldg(reinterpret_cast<const float4*>(x) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (a) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(a + s * C + c));
      v.x *= av.x; v.y *= av.y; v.z *= av.z; v.w *= av.w;
    }
    if (bb) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bb + s * C + c));
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    }
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ x, long long n, int act,
                                                  float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = act_fwd(__ldg(x + i), act);
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                      long long n, int act, float* __restrict__ dx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = __ldg(g + i) * act_grad(__ldg(x + i), act);
}

// out[b, i] = inp[b, i] + x[b, i] * (s ? s[b] : 1)
__global__ void __launch_bounds__(256) add_scaled_kernel(const float* __restrict__ inp, const float* __restrict__ x,
                                                         const float* __restrict__ s, long long per, long long n,
                                                         float* __restrict__ out) {
  const long long nv = n / 4, perv = per / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const float sc = s ? __ldg(s + i / perv) : 1.0f;
    const float4 a = __ldg(reinterpret_cast<const float4*>(inp) + i);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 o;
    o.x = fmaf(v.x, sc, a.x); o.y = fmaf(v.y, sc, a.y); o.z = fmaf(v.z, sc, a.z); o.w = fmaf(v.w, sc, a.w);
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// out[pix, 0:Cn] = x[pix, c0:c0+Cn]      (C, c0, Cn multiples of 4)
__global__ void __launch_bounds__(256) chan_slice_kernel(const float* __restrict__ x, long long npix, int C, int c0,
                                                         int Cn, float* __restrict__ out) {
  const int CV = Cn / 4;
  const long long nv = npix * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / CV;
    const int c = (int)(i - p * CV) * 4;
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(x + p * C + c0 + c));
  }
}

// ------------------------------------------------------------------------- channel reductions
// grid (PB pixel blocks, ceil(C/128) channel tiles, Z samples); block = 32 channel groups (float4)
// x 8 pixel lanes.  part[((z*PB + bx)*NACC + a)*C + c] = this block's sum (double) of term a.
struct MomentsOp {   // (x, x^2), squared in double
  const float* x;
  __device__ __forceinline__ void operator()(long long off, int c, double acc[2][4]) const {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
    const double d[4] = {(double)xv.x, (double)xv.y, (double)xv.z, (double)xv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[0][j] += d[j]; acc[1][j] += d[j] * d[j]; }
  }
};
struct DotOp {       // x * y (y NULL: x)
  const float* x; const float* y;
  __device__ __forceinline__ void operator()(long long off, int c, double acc[1][4]) const {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
    if (y) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y + off));
      acc[0][0] += (double)xv.x * (double)yv.x; acc[0][1] += (double)xv.y * (double)yv.y;
      acc[0][2] += (double)xv.z * (double)yv.z; acc[0][3] += (double)xv.w * (double)yv.w;
    } else {
      acc[0][0] += (double)xv.x; acc[0][1] += (double)xv.y; acc[0][2] += (double)xv.z; acc[0][3] += (double)xv.w;
    }
  }
};
struct BnActBwdOp {  // gu = g * act'(x*a+b) (stored if gu != NULL); terms (gu, gu*x)
  const float* g; const float* x; const float* a; const float* b; float* gu; int act;
  __device__ __forceinline__ void operator()(long long off, int c, double acc[2][4]) const {
    float4 gv = __ldg(reinterpret_cast<const float4*>(g + off));
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
    if (act != ACT_NONE) {
      const float4 aa = __ldg(reinterpret_cast<const float4*>(a + c));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b + c));
      gv.x *= act_grad(fmaf(xv.x, aa.x, bb.x), act); gv.y *= act_grad(fmaf(xv.y, aa.y, bb.y), act);
      gv.z *= act_grad(fmaf(xv.z, aa.z, bb.z), act); gv.w *= act_grad(fmaf(xv.w, aa.w, bb.w), act);
      if (gu) *reinterpret_cast<float4*>(gu + off) = gv;
    }
    acc[0][0] += (double)gv.x; acc[0][1] += (double)gv.y; acc[0][2] += (double)gv.z; acc[0][3] += (double)gv.w;
    acc[1][0] += (double)gv.x * (double)xv.x; acc[1][1] += (double)gv.y * (double)xv.y;
    acc[1][2] += (double)gv.z * (double)xv.z; acc[1][3] += (double)gv.w * (double)xv.w;
  }
};

template <int NACC, class Op>
__global__ void __launch_bounds__(256) chan_reduce_kernel(Op op, int C, long long npix, double* __restrict__ part) {
  __shared__ double s_part[8][32][NACC * 4 + 1];
  const int cg = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + cg * 4;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  const long long zoff = (long long)blockIdx.z * npix;
  double acc[NACC][4];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[a][j] = 0.0;
  if (c < C)
    for (long long p = p0 + pl; p < p1; p += 8) op((zoff + p) * C + c, c, acc);
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) s_part[pl][cg][a * 4 + j] = acc[a][j];
  __syncthreads();
  if (pl == 0 && c < C) {
    const size_t base = ((size_t)blockIdx.z * gridDim.x + blockIdx.x) * NACC;
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double t = s_part[0][cg][a * 4 + j];
        for (int l = 1; l < 8; ++l) t += s_part[l][cg][a * 4 + j];
        part[(base + a) * C + c + j] = t;
      }
  }
}

// The same reduction for C <= 64: a warp holds CG = C/4 channel groups, so the form above would leave 32 - CG lanes
// idle (C = 16: 4 of 32 lanes load anything; measured 0.8 - 1.7 TB/s on the 256x480 layers of the EfficientNet stem and
// first blocks).  Here lane -> (pixel sub-lane, channel group): a warp covers 32 / CG pixels per step and every lane
// loads; the partial sums of the 8 x (32 / CG) pixel lanes are added in a fixed order.
template <int NACC, class Op>
__global__ void __launch_bounds__(256) chan_reduce_small_kernel(Op op, int C, long long npix, double* __restrict__ part) {
  __shared__ double s_part[8][32][NACC * 4 + 1];
  const int lane = threadIdx.x & 31, pl = threadIdx.x >> 5;
  const int CG = C >> 2, nsub = 32 / CG;
  const int cg = lane % CG, sub = lane / CG;
  const int c = cg * 4;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  const long long zoff = (long long)blockIdx.z * npix;
  double acc[NACC][4];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[a][j] = 0.0;
  if (sub < nsub)
    for (long long p = p0 + pl * nsub + sub; p < p1; p += 8 * nsub) op((zoff + p) * C + c, c, acc);
#pragma unroll
  for (int a = 0; a < NACC; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) s_part[pl][lane][a * 4 + j] = acc[a][j];
  __syncthreads();
  if (pl == 0 && lane < CG) {
    const size_t base = ((size_t)blockIdx.z * gridDim.x + blockIdx.x) * NACC;
#pragma unroll
    for (int a = 0; a < NACC; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double t = 0.0;
        for (int l = 0; l < 8; ++l)
          for (int u = 0; u < nsub; ++u) t += s_part[l][u * CG + lane][a * 4 + j];
        part[(base + a) * C + c + j] = t;
      }
  }
}

template <int NACC, class Op>
static void chan_reduce_launch(Op op, int C, long long npix, int PB, int Z, double* part, cudaStream_t st) {
  if (C <= 64 && !getenv("CRESTE_NO_SMALL_REDUCE"))
    chan_reduce_small_kernel<NACC, Op><<<dim3(PB, 1, Z), 256, 0, st>>>(op, C, npix, part);
  else
    chan_reduce_kernel<NACC, Op><<<dim3(PB, ceil_div(C, 128), Z), 256, 0, st>>>(op, C, npix, part);
}

// out[z*n + i] = scale * sum_bx part[(z*PB + bx)*n + i].  32 outputs per block; 8 row lanes per output sum
// rows lane, lane + 8, ... in order, then the 8 lane sums are added in lane order (fixed order => reproducible).
template <class T>
__global__ void __launch_bounds__(256) reduce_parts_kernel(const double* __restrict__ part, int PB, int n, int Z,
                                                           double scale, T* __restrict__ out) {
  __shared__ double s_sum[8][32];
  const int o = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + o;
  const bool ok = i < (long long)n * Z;
  double t = 0.0;
  if (ok) {
    const int z = (int)(i / n), k = (int)(i - (long long)z * n);
    const double* src = part + (size_t)z * PB * n + k;
    for (int b = rl; b < PB; b += 8) t += src[(size_t)b * n];
  }
  s_sum[rl][o] = t;
  __syncthreads();
  if (rl == 0 && ok) {
    double tot = s_sum[0][o];
#pragma unroll
    for (int l = 1; l < 8; ++l) tot += s_sum[l][o];
    out[i] = (T)(tot * scale);
  }
}

static int reduce_pb(long long npix, int Z) {
  long long b = npix / 64;
  if (b < 1) b = 1;
  long long cap = 592 / (Z > 0 ? Z : 1);          // x ceil(C/128) channel tiles: >= 4 CTAs per SM at C >= 128
  if (cap < 4) cap = 4;
  return (int)(b > cap ? cap : b);
}

// ------------------------------------------------------------------------------- depthwise conv
// w [R*R][C] (tap-major), x [N,H,W,C], y [N,P,Q,C]; pad (pt, pl) low, the high side is implied.
template <int R>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         int N, int H, int W, int C, int st, int pt, int pl, int P,
                                                         int Q, float* __restrict__ y) {
  const int CV = C / 4;
  const long long nv = (long long)N * P * Q * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    long long pix = i / CV;
    const int q = (int)(pix % Q); pix /= Q;
    const int p = (int)(pix % P);
    const int n = (int)(pix / P);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int iy = p * st + r - pt;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int ix = q * st + s - pl;
        if (ix < 0 || ix >= W) continue;
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * C + c));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)(r * R + s) * C + c));
        acc.x = fmaf(xv.x, wv.x, acc.x); acc.y = fmaf(xv.y, wv.y, acc.y);
        acc.z = fmaf(xv.z, wv.z, acc.z); acc.w = fmaf(xv.w, wv.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(y)[i] = acc;
  }
}

// x-blocked form (as the eval path's dwconv_xb_kernel, without its BatchNorm / swish / squeeze-excite epilogue): a
// thread produces 4 adjacent outputs of one row for its 4 channels and loads each input column once per filter row
// (3*STRIDE + R float4 loads for 4*R FMAs-by-4 instead of 4*R).  Same (r, s) ascending FMA order per output; a tap
// outside the image adds an exact 0 where dwconv_fwd_kernel skips it.
template <int R, int STRIDE>
__global__ void __launch_bounds__(256) dwconv_fwd_xb_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            int N, int H, int W, int C, int pt, int pl, int P, int Q,
                                                            float* __restrict__ y) {
  constexpr int XB = 4, NC = (XB - 1) * STRIDE + R;
  const unsigned C4 = (unsigned)(C / 4), Qb = (unsigned)((Q + XB - 1) / XB);
  const long long total = (long long)N * P * Qb * C4;              // < 2^32 checked by the launcher
  const float4* wv0 = reinterpret_cast<const float4*>(w);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned u = (unsigned)i;
    const unsigned cg = u % C4; u /= C4;
    const int ox0 = (int)(u % Qb) * XB; u /= Qb;
    const int oy = (int)(u % (unsigned)P);
    const int n = (int)(u / (unsigned)P);
    const int iy0 = oy * STRIDE - pt, ix0 = ox0 * STRIDE - pl;
    const float4* xn = reinterpret_cast<const float4*>(x + (size_t)n * H * W * C) + cg;
    const float4* wv = wv0 + cg;
    float4 acc[XB];
#pragma unroll
    for (int b = 0; b < XB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int iy = iy0 + r;
      if (iy < 0 || iy >= H) continue;
      float4 col[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const int ix = ix0 + c;
        col[c] = (ix >= 0 && ix < W) ? __ldg(xn + ((size_t)iy * W + ix) * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const float4 k = __ldg(wv + (size_t)(r * R + s) * C4);
#pragma unroll
        for (int b = 0; b < XB; ++b) {
          const float4 v = col[b * STRIDE + s];
          acc[b].x = fmaf(v.x, k.x, acc[b].x); acc[b].y = fmaf(v.y, k.y, acc[b].y);
          acc[b].z = fmaf(v.z, k.z, acc[b].z); acc[b].w = fmaf(v.w, k.w, acc[b].w);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < XB; ++b)
      if (ox0 + b < Q) reinterpret_cast<float4*>(y + (((size_t)n * P + oy) * Q + ox0 + b) * C)[cg] = acc[b];
  }
}

// dx[n,iy,ix,c] = sum over taps (r,s) with (iy+pt-r) = p*st, (ix+pl-s) = q*st of w[r,s,c] * g[n,p,q,c]
template <int R>
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                           int N, int H, int W, int C, int st, int pt, int pl, int P,
                                                           int Q, float* __restrict__ dx) {
  const int CV = C / 4;
  const long long nv = (long long)N * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * 4;
    long long pix = i / CV;
    const int ix = (int)(pix % W); pix /= W;
    const int iy = (int)(pix % H);
    const int n = (int)(pix / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int ty = iy + pt - r;
      if (ty < 0 || (ty % st) != 0) continue;
      const int p = ty / st;
      if (p >= P) continue;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        const int tx = ix + pl - s;
        if (tx < 0 || (tx % st) != 0) continue;
        const int q = tx / st;
        if (q >= Q) continue;
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + (((size_t)n * P + p) * Q + q) * C + c));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)(r * R + s) * C + c));
        acc.x = fmaf(gv.x, wv.x, acc.x); acc.y = fmaf(gv.y, wv.y, acc.y);
        acc.z = fmaf(gv.z, wv.z, acc.z); acc.w = fmaf(gv.w, wv.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
  }
}

// Stride 2: only the taps of the right parity contribute (r = (iy + pt) & 1, +2, ...), so the loops visit <= ceil(R/2)^2
// taps instead of testing R*R (runtime `% st` per tap in the generic kernel); same ascending (r, s) FMA order over the
// contributing taps: bit-identical to dwconv_dgrad_kernel.
template <int R>
__global__ void __launch_bounds__(256) dwconv_dgrad_s2_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                              int N, int H, int W, int C, int pt, int pl, int P,
                                                              int Q, float* __restrict__ dx) {
  const int CV = C / 4;
  const long long nv = (long long)N * H * W * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;                       // nv < 2^32 checked by the launcher
    const int c = (int)(iu % (unsigned)CV) * 4;
    unsigned pix = iu / (unsigned)CV;
    const int ix = (int)(pix % (unsigned)W); pix /= (unsigned)W;
    const int iy = (int)(pix % (unsigned)H);
    const int n = (int)(pix / (unsigned)H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int r0 = (iy + pt) & 1, s0 = (ix + pl) & 1;
#pragma unroll
    for (int rr = 0; rr < (R + 1) / 2; ++rr) {
      const int r = r0 + 2 * rr;
      const int ty = iy + pt - r;
      if (r >= R || ty < 0) continue;
      const int p = ty >> 1;
      if (p >= P) continue;
#pragma unroll
      for (int ss = 0; ss < (R + 1) / 2; ++ss) {
        const int sx = s0 + 2 * ss;
        const int tx = ix + pl - sx;
        if (sx >= R || tx < 0) continue;
        const int q = tx >> 1;
        if (q >= Q) continue;
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + (((size_t)n * P + p) * Q + q) * C + c));
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)(r * R + sx) * C + c));
        acc.x = fmaf(gv.x, wv.x, acc.x); acc.y = fmaf(gv.y, wv.y, acc.y);
        acc.z = fmaf(gv.z, wv.z, acc.z); acc.w = fmaf(gv.w, wv.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
  }
}

// Stride-1 data gradient from a shared-memory tile of g (with halo, zero outside the output grid): the same (r, s)
// ascending FMA order per element as dwconv_dgrad_kernel, so the results are bit-identical to it; 4 adjacent outputs
// per thread, taps without bounds checks.
template <int R>
__global__ void __launch_bounds__(256) dwconv_dgrad_tile_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                                int H, int W, int C, int pt, int pl, int P, int Q,
                                                                int tiles_x, float* __restrict__ dx) {
  constexpr int TH = 8, TW = 32, XB = 4, CG = 8;
  constexpr int IH = TH + R - 1, IW = TW + R - 1, NC = XB + R - 1;
  constexpr int UNITS = TH * (TW / XB);
  extern __shared__ __align__(16) float4 dg_smem[];
  float4* tile = dg_smem;                          // g rows iy0 + pt - (R-1) .. , cols ix0 + pl - (R-1) ..
  float4* wsm = dg_smem + IH * IW * CG;            // [R*R][CG]
  const int C4 = C / 4;
  const int n = blockIdx.y;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int iy0 = ty * TH, ix0 = tx * TW;
  const int cg0 = blockIdx.z * CG;
  const int cgs = min(CG, C4 - cg0);
  const int tid = threadIdx.x;
  const float4* gn = reinterpret_cast<const float4*>(g + (size_t)n * P * Q * C);
  const int p0 = iy0 + pt - (R - 1), q0 = ix0 + pl - (R - 1);
  for (int i = tid; i < IH * IW * CG; i += 256) {
    const int c = i % CG, e = i / CG;
    const int q = q0 + e % IW, p = p0 + e / IW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cgs && p >= 0 && p < P && q >= 0 && q < Q) v = __ldg(gn + ((size_t)p * Q + q) * C4 + cg0 + c);
    tile[i] = v;
  }
  for (int i = tid; i < R * R * CG; i += 256) {
    const int c = i % CG, t = i / CG;
    wsm[i] = c < cgs ? __ldg(reinterpret_cast<const float4*>(w) + (size_t)t * C4 + cg0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int c = tid % CG, ul = tid / CG;
  if (c >= cgs) return;
#pragma unroll 1
  for (int u = ul; u < UNITS; u += 32) {
    const int ly = u / (TW / XB), lx = (u - ly * (TW / XB)) * XB;
    const int iy = iy0 + ly;
    if (iy >= H || ix0 + lx >= W) continue;
    float4 acc[XB];
#pragma unroll
    for (int b = 0; b < XB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      // tap (r, s) of output (ly, lx + b) reads g at tile row ly + (R-1) - r, tile col lx + b + (R-1) - s
      const float4* row = tile + ((size_t)(ly + R - 1 - r) * IW + lx) * CG + c;
      float4 col[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) col[k] = row[(size_t)k * CG];
#pragma unroll
      for (int s_ = 0; s_ < R; ++s_) {
        const float4 k4 = wsm[(r * R + s_) * CG + c];
#pragma unroll
        for (int b = 0; b < XB; ++b) {
          const float4 v = col[b + R - 1 - s_];
          acc[b].x = fmaf(v.x, k4.x, acc[b].x); acc[b].y = fmaf(v.y, k4.y, acc[b].y);
          acc[b].z = fmaf(v.z, k4.z, acc[b].z); acc[b].w = fmaf(v.w, k4.w, acc[b].w);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < XB; ++b)
      if (ix0 + lx + b < W)
        reinterpret_cast<float4*>(dx + ((size_t)(n * (size_t)H + iy) * W + ix0 + lx + b) * C)[cg0 + c] = acc[b];
  }
}

// part[bx][tap][c] = sum over the block's output pixels of g[pix,c] * x[shifted pix,c]; per-thread
// fp32 partial sums over <= a few hundred pixels, the cross-lane / cross-block stages in double.
template <int R>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           int N, int H, int W, int C, int st, int pt, int pl, int P,
                                                           int Q, double* __restrict__ part) {
  __shared__ float s_red[8][32][4];
  const int cg = threadIdx.x & 31, pl_ = threadIdx.x >> 5;
  const int c = blockIdx.y * 128 + cg * 4;
  const long long npix = (long long)N * P * Q;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p0 = (long long)blockIdx.x * per;
  const long long p1 = p0 + per < npix ? p0 + per : npix;
  float4 acc[R * R];
#pragma unroll
  for (int t = 0; t < R * R; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    for (long long pix = p0 + pl_; pix < p1; pix += 8) {
      const int q = (int)(pix % Q);
      const long long t2 = pix / Q;
      const int p = (int)(t2 % P);
      const int n = (int)(t2 / P);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + (size_t)pix * C + c));
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int iy = p * st + r - pt;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int s = 0; s < R; ++s) {
          const int ix = q * st + s - pl;
          if (ix < 0 || ix >= W) continue;
          const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * C + c));
          float4& a = acc[r * R + s];
          a.x = fmaf(gv.x, xv.x, a.x); a.y = fmaf(gv.y, xv.y, a.y);
          a.z = fmaf(gv.z, xv.z, a.z); a.w = fmaf(gv.w, xv.w, a.w);
        }
      }
    }
  }
#pragma unroll 1
  for (int t = 0; t < R * R; ++t) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    // static indexing of acc[] (registers): select by an unrolled compare chain
#pragma unroll
    for (int u = 0; u < R * R; ++u)
      if (u == t) a = acc[u];
    s_red[pl_][cg][0] = a.x; s_red[pl_][cg][1] = a.y; s_red[pl_][cg][2] = a.z; s_red[pl_][cg][3] = a.w;
    __syncthreads();
    if (pl_ == 0 && c < C) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double tot = (double)s_red[0][cg][j];
        for (int l = 1; l < 8; ++l) tot += (double)s_red[l][cg][j];
        part[((size_t)blockIdx.x * (R * R) + t) * C + c + j] = tot;
      }
    }
    __syncthreads();
  }
}

// Tiled form (round 2b): the kernel above keeps R*R float4 accumulators per thread (100+ registers at R = 5, two CTAs
// per SM), re-derives three 64-bit divisions per pixel, reads every tap from global memory and idles most lanes of the
// second channel block when C is not a multiple of 128 -- 650 us per 5x5 layer at B = 16, 15x its HBM floor.  Here a
// CTA stages the x tile (with halo, zero outside the image) and the g tile of 8 x TW outputs x 32 channels in shared
// memory; thread (tap, 4-channel group, row lane) accumulates g * x over its rows of the tile in fp32 (<= 256 pixels),
// the row lanes are added in a fixed order and the tile's [R*R][32] partial goes out in double;
// reduce_parts_kernel adds the tiles in its fixed order.
template <int R, int STRIDE>
__global__ void __launch_bounds__(256) dwconv_wgrad_tile_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                int H, int W, int C, int pt, int pl, int P, int Q,
                                                                int tiles_x, double* __restrict__ part) {
  constexpr int TH = 8, TW = STRIDE == 1 ? 32 : 16, CG = 8;
  constexpr int IH = (TH - 1) * STRIDE + R, IW = (TW - 1) * STRIDE + R;
  constexpr int RR = R * R;
  constexpr int PG = 32 / RR >= 1 ? 32 / RR : 1;      // row lanes per (tap, channel group): 3 at R = 3, 1 at R = 5
  extern __shared__ __align__(16) float4 wg_smem[];
  float4* xt = wg_smem;                                // [IH][IW][CG]
  float4* gt = wg_smem + IH * IW * CG;                 // [TH][TW][CG]
  float4* red = gt + TH * TW * CG;                     // [PG][RR][CG]
  const int C4 = C / 4;
  const int n = blockIdx.y;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int oy0 = ty * TH, ox0 = tx * TW;
  const int cg0 = blockIdx.z * CG;
  const int cgs = min(CG, C4 - cg0);
  const int tid = threadIdx.x;
  const float4* xn = reinterpret_cast<const float4*>(x + (size_t)n * H * W * C);
  const float4* gn = reinterpret_cast<const float4*>(g + (size_t)n * P * Q * C);
  const int iy0 = oy0 * STRIDE - pt, ix0 = ox0 * STRIDE - pl;
  for (int i = tid; i < IH * IW * CG; i += 256) {
    const int c = i % CG, p = i / CG;
    const int ix = ix0 + p % IW, iy = iy0 + p / IW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cgs && iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xn + ((size_t)iy * W + ix) * C4 + cg0 + c);
    xt[i] = v;
  }
  for (int i = tid; i < TH * TW * CG; i += 256) {
    const int c = i % CG, p = i / CG;
    const int ox = ox0 + p % TW, oy = oy0 + p / TW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cgs && oy < P && ox < Q) v = __ldg(gn + ((size_t)oy * Q + ox) * C4 + cg0 + c);
    gt[i] = v;
  }
  __syncthreads();
  const int c = tid % CG, t2 = tid / CG;               // t2 = pg * RR + tap
  const int tap = t2 % RR, pg = t2 / RR;
  if (pg < PG) {
    const int r = tap / R, s_ = tap - r * R;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ly = pg; ly < TH; ly += PG) {
      const float4* xr = xt + ((size_t)(ly * STRIDE + r) * IW + s_) * CG + c;
      const float4* gr = gt + (size_t)ly * TW * CG + c;
#pragma unroll 8
      for (int lx = 0; lx < TW; ++lx) {
        const float4 gv = gr[lx * CG], xv = xr[(size_t)lx * STRIDE * CG];
        acc.x = fmaf(gv.x, xv.x, acc.x); acc.y = fmaf(gv.y, xv.y, acc.y);
        acc.z = fmaf(gv.z, xv.z, acc.z); acc.w = fmaf(gv.w, xv.w, acc.w);
      }
    }
    red[(pg * RR + tap) * CG + c] = acc;
  }
  __syncthreads();
  if (tid < RR * CG) {
    const int cc = tid % CG, tp = tid / CG;
    if (cc < cgs) {
      const float4 a0 = red[tp * CG + cc];
      double d0 = (double)a0.x, d1 = (double)a0.y, d2 = (double)a0.z, d3 = (double)a0.w;
      for (int l = 1; l < PG; ++l) {
        const float4 a = red[(l * RR + tp) * CG + cc];
        d0 += (double)a.x; d1 += (double)a.y; d2 += (double)a.z; d3 += (double)a.w;
      }
      double* dst = part + (((size_t)n * gridDim.x + blockIdx.x) * RR + tp) * C + (size_t)(cg0 + cc) * 4;
      dst[0] = d0; dst[1] = d1; dst[2] = d2; dst[3] = d3;
    }
  }
}

// ------------------------------------------------------------------- strided dense wgrad (C == 4)
// thread -> (tap, k); part[bx][(tap*4 + c)*K + k] = sum over the block's output pixels.
// A CTA walks tiles of WG_TQ output pixels of one output row: the tile's g [WG_TQ][K] and the R input rows it touches
// ([R][(WG_TQ-1)*st + S] float4, zero outside the image) are staged in shared memory with coalesced loads, then every
// thread runs the pixel loop out of shared memory (g: consecutive k, conflict-free; x: one float4 broadcast per warp).
// fp32 accumulation inside a tile, double across tiles.  The round-2b form read g and x from global memory inside a
// per-thread pixel loop with two 64-bit divisions per pixel: 2.1 ms for the 512x960 stem at B = 16; this one is
// bound by the 380 MB it reads.
constexpr int WG_TQ = 64;
__global__ void __launch_bounds__(1024) wgrad_strided_c4_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ g, int N, int H, int W,
                                                                int K, int R, int S, int st, int pt, int pl, int P,
                                                                int Q, int tiles_q, long long ntiles,
                                                                double* __restrict__ part) {
  extern __shared__ float4 wg_smem[];
  const int XW = (WG_TQ - 1) * st + S;                 // input columns a tile touches
  float4* sx = wg_smem;                                // [R][XW]
  float* sg = reinterpret_cast<float*>(wg_smem + R * XW);      // [WG_TQ][K]
  const int t = threadIdx.x;
  const int tap = t / K, k = t - tap * K;
  const bool live = tap < R * S;
  const int r = live ? tap / S : 0, s = live ? tap - r * S : 0;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tq = (int)(tile % tiles_q);
    const long long row = tile / tiles_q;              // n * P + p
    const int p = (int)(row % P), n = (int)(row / P);
    const int q0 = tq * WG_TQ;
    const int nq = min(WG_TQ, Q - q0);
    __syncthreads();                                   // the previous tile's readers are done
    for (int i = t; i < R * XW; i += blockDim.x) {
      const int rr = i / XW, cc = i - rr * XW;
      const int iy = p * st + rr - pt, ix = q0 * st + cc - pl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * 4));
      sx[i] = v;
    }
    const float* gsrc = g + ((size_t)row * Q + q0) * K;
    const int ng = nq * K;                             // contiguous in global memory
    if ((K & 3) == 0) {
      for (int i = t; i < ng / 4; i += blockDim.x)
        reinterpret_cast<float4*>(sg)[i] = __ldg(reinterpret_cast<const float4*>(gsrc) + i);
    } else {
      for (int i = t; i < ng; i += blockDim.x) sg[i] = __ldg(gsrc + i);
    }
    __syncthreads();
    if (live) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* xr = sx + r * XW + s;
#pragma unroll 4
      for (int j = 0; j < nq; ++j) {
        const float gv = sg[j * K + k];
        const float4 xv = xr[j * st];
        acc.x = fmaf(gv, xv.x, acc.x); acc.y = fmaf(gv, xv.y, acc.y);
        acc.z = fmaf(gv, xv.z, acc.z); acc.w = fmaf(gv, xv.w, acc.w);
      }
      a0 += (double)acc.x; a1 += (double)acc.y; a2 += (double)acc.z; a3 += (double)acc.w;
    }
  }
  if (live) {
    double* dst = part + (size_t)blockIdx.x * (R * S * 4 * K);
    dst[(tap * 4 + 0) * K + k] = a0; dst[(tap * 4 + 1) * K + k] = a1;
    dst[(tap * 4 + 2) * K + k] = a2; dst[(tap * 4 + 3) * K + k] = a3;
  }
}

// ---------------------------------------------------------------- BatchNorm per-channel algebra
// forward: batch moments -> (a, b) of y = x*a + b, saved (mean, inv) and the running-stat update of
// F.batch_norm(training=True) (momentum form; unbiased variance).  One launch instead of ~15 [C]-sized ops.
__global__ void __launch_bounds__(256) bn_fwd_finalize_kernel(const double* __restrict__ st, const float* __restrict__ weight,
                                                              const float* __restrict__ bias, int C, double M, double eps,
                                                              float momentum, float* __restrict__ running_mean,
                                                              float* __restrict__ running_var, float* __restrict__ ab,
                                                              double* __restrict__ mi) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = st[c] / M;
  double var = st[C + c] / M - mean * mean;
  var = var > 0.0 ? var : 0.0;
  if (running_mean) {
    const double unb = var * (M / (M > 1.0 ? M - 1.0 : 1.0));
    running_mean[c] = running_mean[c] * (1.0f - momentum) + momentum * (float)mean;
    running_var[c] = running_var[c] * (1.0f - momentum) + momentum * (float)unb;
  }
  const double inv = 1.0 / sqrt(var + eps);
  const double a = inv * (weight ? (double)weight[c] : 1.0);
  const double b = (bias ? (double)bias[c] : 0.0) - mean * a;
  ab[c] = (float)a;
  ab[C + c] = (float)b;
  mi[c] = mean;
  mi[C + c] = inv;
}

// backward: (sum gu, sum gu*x) -> dgamma, dbeta and the (q, r) of dx = gu*a + x*q + r
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ ab,
                                                              const double* __restrict__ mi, int C, double M,
                                                              float* __restrict__ out4) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = sums[c], s2 = sums[C + c], mean = mi[c], inv = mi[C + c], a = (double)ab[c];
  const double dgamma = inv * (s2 - mean * s1);
  const double q = -a * inv * dgamma / M;
  const double r = -a * s1 / M - q * mean;
  out4[c] = (float)dgamma;
  out4[C + c] = (float)s1;
  out4[2 * C + c] = (float)q;
  out4[3 * C + c] = (float)r;
}

// ------------------------------------------------- 1x1 wgrad over a handful of rows (squeeze-excite)
// dw[c][k] = sum_p x[p][c] * g[p][k], p < npix (= batch size: the SE convs act on [B,1,1,C] vectors)
__global__ void __launch_bounds__(256) wgrad_rows_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                         int npix, int C, int K, float* __restrict__ dw) {
  const long long n = (long long)C * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / K), k = (int)(i - (long long)c * K);
    float acc = 0.f;
    for (int p = 0; p < npix; ++p) acc = fmaf(__ldg(x + (size_t)p * C + c), __ldg(g + (size_t)p * K + k), acc);
    dw[i] = acc;
  }
}

// ----------------------------------------------------------------------------- loss gradients
// dlogits[n,k,p] = scale * (softmax_k(logits[n,:,p]) - [k == bin]) on valid pixels, 0 elsewhere
// (CrossEntropyDepth, loss_utils.py:477-527; bin_depths 'UD', depth_utils.py:346-383).  NCHW.
__global__ void __launch_bounds__(256) ce_depth_bwd_kernel(const float* __restrict__ logits,
                                                           const float* __restrict__ label_mm, int N, int D,
                                                           long long HW, float dmin, float bin_size,
                                                           const float* __restrict__ scale_dev,
                                                           float* __restrict__ dlogits) {
  const float scale = __ldg(scale_dev);
  const long long total = (long long)N * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, p = i - n * HW;
    const float d = __ldg(label_mm + i);
    const float idxf = __fdiv_rn(__fsub_rn(d, dmin), bin_size);
    const bool bad = (idxf < 0.0f) || (idxf > (float)D) || !isfinite(idxf);
    const int bin = bad ? D : (int)idxf;
    const float* src = logits + (size_t)n * D * HW + p;
    float* dst = dlogits + (size_t)n * D * HW + p;
    if (bin == D) {
      for (int k = 0; k < D; ++k) dst[(size_t)k * HW] = 0.f;
      continue;
    }
    float m = -INFINITY;
    for (int k = 0; k < D; ++k) m = fmaxf(m, __ldg(src + (size_t)k * HW));
    float ssum = 0.f;
    for (int k = 0; k < D; ++k) ssum += expf(__ldg(src + (size_t)k * HW) - m);
    const float inv = 1.0f / ssum;
    for (int k = 0; k < D; ++k) {
      const float pk = expf(__ldg(src + (size_t)k * HW) - m) * inv;
      dst[(size_t)k * HW] = scale * (pk - (k == bin ? 1.0f : 0.0f));
    }
  }
}

// dpred = scale * (pred - gt) where gt is not +-inf, 0 elsewhere   (MSELoss, loss_utils.py:606-647)
__global__ void __launch_bounds__(256) masked_mse_bwd_kernel(const float* __restrict__ pred,
                                                             const float* __restrict__ gt, long long n,
                                                             const float* __restrict__ scale_dev,
                                                             float* __restrict__ dpred) {
  const float scale = __ldg(scale_dev);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float t = __ldg(gt + i);
    dpred[i] = isinf(t) ? 0.f : scale * (__ldg(pred + i) - t);
  }
}

}  // namespace creste

using namespace creste;

#define ELT_GRID(nvec) grid_cap2((nvec), 256, 148 * 16)

extern "C" size_t creste_chan_reduce_workspace_bytes(long long npix, int C, int nacc, int Z) {
  return (size_t)reduce_pb(npix, Z) * (Z > 0 ? Z : 1) * nacc * C * sizeof(double);
}

extern "C" int creste_chan_moments(const float* x, long long npix, int C, double* out2, void* ws, size_t ws_bytes,
                                   void* stream) {
  CRESTE_CHECK_ARG(x && out2 && ws && npix > 0 && C > 0 && C % 4 == 0, "creste_chan_moments: bad args");
  CRESTE_CHECK_ARG(ws_bytes >= creste_chan_reduce_workspace_bytes(npix, C, 2, 1), "creste_chan_moments: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int PB = reduce_pb(npix, 1);
  chan_reduce_launch<2>(MomentsOp{x}, C, npix, PB, 1, (double*)ws, st);
  int rc = launch_check("chan_reduce_kernel<moments>");
  if (rc) return rc;
  reduce_parts_kernel<double><<<ceil_div(2 * C, 32), 256, 0, st>>>((const double*)ws, PB, 2 * C, 1, 1.0, out2);
  return launch_check("reduce_parts_kernel");
}

extern "C" int creste_chan_affine_act(const float* x, const float* a, const float* b, long long npix, int C, int act,
                                      float* y, void* stream) {
  CRESTE_CHECK_ARG(x && y && npix > 0 && C > 0 && C % 4 == 0 && act >= 0 && act <= 2, "creste_chan_affine_act: bad args");
  const long long n = npix * C;
  chan_affine_act_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(x, a, b, C, n, act, y, nullptr);
  return launch_check("chan_affine_act_kernel");
}

extern "C" int creste_chan_affine_act_amax(const float* x, const float* a, const float* b, long long npix, int C,
                                           int act, float* y, float* amax_out, void* stream) {
  CRESTE_CHECK_ARG(x && y && amax_out && npix > 0 && C > 0 && C % 4 == 0 && act >= 0 && act <= 2,
                   "creste_chan_affine_act_amax: bad args");
  const long long n = npix * C;
  CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, 4, (cudaStream_t)stream));
  chan_affine_act_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(x, a, b, C, n, act, y, (unsigned*)amax_out);
  return launch_check("chan_affine_act_kernel");
}

extern "C" int creste_bn_act_bwd(const float* g, const float* x, const float* a, const float* b, long long npix, int C,
                                 int act, float* gu, double* sums2, void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(g && x && sums2 && ws && npix > 0 && C > 0 && C % 4 == 0 && act >= 0 && act <= 2,
                   "creste_bn_act_bwd: bad args");
  CRESTE_CHECK_ARG(act == ACT_NONE || (a && b), "creste_bn_act_bwd: act needs a and b");
  CRESTE_CHECK_ARG(ws_bytes >= creste_chan_reduce_workspace_bytes(npix, C, 2, 1), "creste_bn_act_bwd: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int PB = reduce_pb(npix, 1);
  chan_reduce_launch<2>(BnActBwdOp{g, x, a, b, gu, act}, C, npix, PB, 1, (double*)ws, st);
  int rc = launch_check("chan_reduce_kernel<bn_act_bwd>");
  if (rc) return rc;
  reduce_parts_kernel<double><<<ceil_div(2 * C, 32), 256, 0, st>>>((const double*)ws, PB, 2 * C, 1, 1.0, sums2);
  return launch_check("reduce_parts_kernel");
}

extern "C" int creste_chan_axpby(const float* u, const float* x, const float* p, const float* q, const float* r,
                                 long long npix, int C, float* out, void* stream) {
  CRESTE_CHECK_ARG(u && x && p && q && r && out && npix > 0 && C > 0 && C % 4 == 0, "creste_chan_axpby: bad args");
  const long long n = npix * C;
  chan_axpby_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(u, x, p, q, r, C, n, out, nullptr);
  return launch_check("chan_axpby_kernel");
}

extern "C" int creste_chan_axpby_act(const float* g, const float* x, const float* a, const float* b, int act,
                                     const float* p, const float* q, const float* r, long long npix, int C, float* out,
                                     float* amax_out, void* stream) {
  CRESTE_CHECK_ARG(g && x && a && b && p && q && r && out && npix > 0 && C > 0 && C % 4 == 0 && act >= 0 && act <= 2,
                   "creste_chan_axpby_act: bad args");
  const long long n = npix * C;
  if (amax_out) CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, 4, (cudaStream_t)stream));
  chan_axpby_act_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(g, x, a, b, act, p, q, r, C, n, out,
                                                                           (unsigned*)amax_out);
  return launch_check("chan_axpby_act_kernel");
}

extern "C" int creste_chan_axpby_amax(const float* u, const float* x, const float* p, const float* q, const float* r,
                                      long long npix, int C, float* out, float* amax_out, void* stream) {
  CRESTE_CHECK_ARG(u && x && p && q && r && out && amax_out && npix > 0 && C > 0 && C % 4 == 0,
                   "creste_chan_axpby_amax: bad args");
  const long long n = npix * C;
  CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, 4, (cudaStream_t)stream));
  chan_axpby_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(u, x, p, q, r, C, n, out, (unsigned*)amax_out);
  return launch_check("chan_axpby_kernel");
}

static int dw_geom_ok(int N, int H, int W, int C, int R, int st, int P, int Q) {
  return N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && (R == 3 || R == 5) && (st == 1 || st == 2) && P > 0 && Q > 0;
}

extern "C" int creste_dwconv_fwd(const float* x, const float* w, int N, int H, int W, int C, int R, int stride,
                                 int pad_t, int pad_l, int P, int Q, float* y, void* stream) {
  CRESTE_CHECK_ARG(x && w && y && dw_geom_ok(N, H, W, C, R, stride, P, Q), "creste_dwconv_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long units = (long long)N * P * ((Q + 3) / 4) * (C / 4);
  if ((stride == 1 || stride == 2) && units < (1LL << 32) && !getenv("CRESTE_NO_DWFWD_XB")) {
    const int xgrid = ELT_GRID(units);
    if (R == 3 && stride == 1) dwconv_fwd_xb_kernel<3, 1><<<xgrid, 256, 0, st>>>(x, w, N, H, W, C, pad_t, pad_l, P, Q, y);
    else if (R == 3) dwconv_fwd_xb_kernel<3, 2><<<xgrid, 256, 0, st>>>(x, w, N, H, W, C, pad_t, pad_l, P, Q, y);
    else if (stride == 1) dwconv_fwd_xb_kernel<5, 1><<<xgrid, 256, 0, st>>>(x, w, N, H, W, C, pad_t, pad_l, P, Q, y);
    else dwconv_fwd_xb_kernel<5, 2><<<xgrid, 256, 0, st>>>(x, w, N, H, W, C, pad_t, pad_l, P, Q, y);
    return launch_check("dwconv_fwd_xb_kernel");
  }
  const int grid = ELT_GRID((long long)N * P * Q * (C / 4));
  if (R == 3) dwconv_fwd_kernel<3><<<grid, 256, 0, st>>>(x, w, N, H, W, C, stride, pad_t, pad_l, P, Q, y);
  else dwconv_fwd_kernel<5><<<grid, 256, 0, st>>>(x, w, N, H, W, C, stride, pad_t, pad_l, P, Q, y);
  return launch_check("dwconv_fwd_kernel");
}

extern "C" int creste_dwconv_dgrad(const float* g, const float* w, int N, int H, int W, int C, int R, int stride,
                                   int pad_t, int pad_l, int P, int Q, float* dx, void* stream) {
  CRESTE_CHECK_ARG(g && w && dx && dw_geom_ok(N, H, W, C, R, stride, P, Q), "creste_dwconv_dgrad: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (stride == 1 && !getenv("CRESTE_NO_DWDGRAD_TILE")) {
    const int tiles_x = ceil_div(W, 32), tiles = tiles_x * ceil_div(H, 8);
    const size_t smem = ((size_t)(8 + R - 1) * (32 + R - 1) * 8 + (size_t)R * R * 8) * sizeof(float4);
    const dim3 tgrid(tiles, N, ceil_div(C / 4, 8));
    if (R == 3) {
      CRESTE_CUDA(cudaFuncSetAttribute(dwconv_dgrad_tile_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dwconv_dgrad_tile_kernel<3><<<tgrid, 256, smem, st>>>(g, w, H, W, C, pad_t, pad_l, P, Q, tiles_x, dx);
    } else {
      CRESTE_CUDA(cudaFuncSetAttribute(dwconv_dgrad_tile_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      dwconv_dgrad_tile_kernel<5><<<tgrid, 256, smem, st>>>(g, w, H, W, C, pad_t, pad_l, P, Q, tiles_x, dx);
    }
    return launch_check("dwconv_dgrad_tile_kernel");
  }
  const int grid = ELT_GRID((long long)N * H * W * (C / 4));
  if (stride == 2 && (long long)N * H * W * (C / 4) < (1LL << 32) && !getenv("CRESTE_NO_DWDGRAD_S2")) {
    if (R == 3) dwconv_dgrad_s2_kernel<3><<<grid, 256, 0, st>>>(g, w, N, H, W, C, pad_t, pad_l, P, Q, dx);
    else dwconv_dgrad_s2_kernel<5><<<grid, 256, 0, st>>>(g, w, N, H, W, C, pad_t, pad_l, P, Q, dx);
    return launch_check("dwconv_dgrad_s2_kernel");
  }
  if (R == 3) dwconv_dgrad_kernel<3><<<grid, 256, 0, st>>>(g, w, N, H, W, C, stride, pad_t, pad_l, P, Q, dx);
  else dwconv_dgrad_kernel<5><<<grid, 256, 0, st>>>(g, w, N, H, W, C, stride, pad_t, pad_l, P, Q, dx);
  return launch_check("dwconv_dgrad_kernel");
}

extern "C" size_t creste_dwconv_wgrad_workspace_bytes(int N, int C, int R, int P, int Q) {
  // the tiled kernel writes one [R*R][C] partial per 8 x 16 (stride 2) / 8 x 32 (stride 1) output tile: size for 8 x 16
  const size_t tiled = (size_t)N * ceil_div(P, 8) * ceil_div(Q, 16) * R * R * C * sizeof(double);
  const size_t flat = (size_t)reduce_pb((long long)N * P * Q, 1) * R * R * C * sizeof(double);
  return tiled > flat ? tiled : flat;
}

extern "C" int creste_dwconv_wgrad(const float* x, const float* g, int N, int H, int W, int C, int R, int stride,
                                   int pad_t, int pad_l, int P, int Q, float* dw, void* ws, size_t ws_bytes,
                                   void* stream) {
  CRESTE_CHECK_ARG(x && g && dw && ws && dw_geom_ok(N, H, W, C, R, stride, P, Q), "creste_dwconv_wgrad: bad args");
  CRESTE_CHECK_ARG(ws_bytes >= creste_dwconv_wgrad_workspace_bytes(N, C, R, P, Q), "creste_dwconv_wgrad: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  if (!getenv("CRESTE_NO_DWWGRAD_TILE")) {
    const int TW = stride == 1 ? 32 : 16;
    const int tiles_x = ceil_div(Q, TW), tiles = tiles_x * ceil_div(P, 8);
    const int IH = 7 * stride + R, IW = (TW - 1) * stride + R;
    const int PG = 32 / (R * R) >= 1 ? 32 / (R * R) : 1;
    const size_t smem = ((size_t)IH * IW * 8 + (size_t)8 * TW * 8 + (size_t)PG * R * R * 8) * sizeof(float4);
    const dim3 tgrid(tiles, N, ceil_div(C / 4, 8));
#define CRESTE_DWWG(RR_, SS_)                                                                                          \
    do {                                                                                                               \
      CRESTE_CUDA(cudaFuncSetAttribute(dwconv_wgrad_tile_kernel<RR_, SS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                                    \
      dwconv_wgrad_tile_kernel<RR_, SS_><<<tgrid, 256, smem, st>>>(x, g, H, W, C, pad_t, pad_l, P, Q, tiles_x,           \
                                                                  (double*)ws);                                        \
    } while (0)
    if (R == 3 && stride == 1) CRESTE_DWWG(3, 1);
    else if (R == 3) CRESTE_DWWG(3, 2);
    else if (stride == 1) CRESTE_DWWG(5, 1);
    else CRESTE_DWWG(5, 2);
#undef CRESTE_DWWG
    int rc = launch_check("dwconv_wgrad_tile_kernel");
    if (rc) return rc;
    reduce_parts_kernel<float><<<ceil_div(R * R * C, 32), 256, 0, st>>>((const double*)ws, N * tiles, R * R * C, 1, 1.0, dw);
    return launch_check("reduce_parts_kernel");
  }
  const int PB = reduce_pb((long long)N * P * Q, 1);
  const dim3 grid(PB, ceil_div(C, 128), 1);
  if (R == 3) dwconv_wgrad_kernel<3><<<grid, 256, 0, st>>>(x, g, N, H, W, C, stride, pad_t, pad_l, P, Q, (double*)ws);
  else dwconv_wgrad_kernel<5><<<grid, 256, 0, st>>>(x, g, N, H, W, C, stride, pad_t, pad_l, P, Q, (double*)ws);
  int rc = launch_check("dwconv_wgrad_kernel");
  if (rc) return rc;
  reduce_parts_kernel<float><<<ceil_div(R * R * C, 32), 256, 0, st>>>((const double*)ws, PB, R * R * C, 1, 1.0, dw);
  return launch_check("reduce_parts_kernel");
}

extern "C" int creste_sample_dot(const float* x, const float* y, int B, long long HW, int C, float scale, float* out,
                                 void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(x && out && ws && B > 0 && HW > 0 && C > 0 && C % 4 == 0, "creste_sample_dot: bad args");
  CRESTE_CHECK_ARG(ws_bytes >= creste_chan_reduce_workspace_bytes(HW, C, 1, B), "creste_sample_dot: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int PB = reduce_pb(HW, B);
  chan_reduce_launch<1>(DotOp{x, y}, C, HW, PB, B, (double*)ws, st);
  int rc = launch_check("chan_reduce_kernel<sample_dot>");
  if (rc) return rc;
  reduce_parts_kernel<float><<<ceil_div(B * C, 32), 256, 0, st>>>((const double*)ws, PB, C, B, (double)scale, out);
  return launch_check("reduce_parts_kernel");
}

extern "C" int creste_sample_affine(const float* x, const float* a, const float* b, int B, long long HW, int C,
                                    float* out, void* stream) {
  CRESTE_CHECK_ARG(out && (x || b) && B > 0 && HW > 0 && C > 0 && C % 4 == 0, "creste_sample_affine: bad args");
  const long long n = (long long)B * HW * C;
  sample_affine_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(x, a, b, HW, C, n, out);
  return launch_check("sample_affine_kernel");
}

extern "C" int creste_act(const float* x, long long n, int act, float* y, void* stream) {
  CRESTE_CHECK_ARG(x && y && n > 0 && act >= 0 && act <= 3, "creste_act: bad args");
  act_kernel<<<ELT_GRID(n), 256, 0, (cudaStream_t)stream>>>(x, n, act, y);
  return launch_check("act_kernel");
}

extern "C" int creste_act_bwd(const float* g, const float* x, long long n, int act, float* dx, void* stream) {
  CRESTE_CHECK_ARG(g && x && dx && n > 0 && act >= 0 && act <= 3, "creste_act_bwd: bad args");
  act_bwd_kernel<<<ELT_GRID(n), 256, 0, (cudaStream_t)stream>>>(g, x, n, act, dx);
  return launch_check("act_bwd_kernel");
}

extern "C" int creste_add_scaled(const float* inp, const float* x, const float* s, int B, long long per, float* out,
                                 void* stream) {
  CRESTE_CHECK_ARG(inp && x && out && B > 0 && per > 0 && per % 4 == 0, "creste_add_scaled: bad args");
  const long long n = (long long)B * per;
  add_scaled_kernel<<<ELT_GRID(n / 4), 256, 0, (cudaStream_t)stream>>>(inp, x, s, per, n, out);
  return launch_check("add_scaled_kernel");
}

extern "C" int creste_chan_slice(const float* x, long long npix, int C, int c0, int Cn, float* out, void* stream) {
  CRESTE_CHECK_ARG(x && out && npix > 0 && C % 4 == 0 && c0 % 4 == 0 && Cn % 4 == 0 && c0 >= 0 && Cn > 0 &&
                       c0 + Cn <= C, "creste_chan_slice: bad args");
  chan_slice_kernel<<<ELT_GRID(npix * (Cn / 4)), 256, 0, (cudaStream_t)stream>>>(x, npix, C, c0, Cn, out);
  return launch_check("chan_slice_kernel");
}

static int wgrad_strided_pb(long long npix) {
  long long b = npix / WG_TQ;               // one tile of WG_TQ output pixels at least
  if (b < 1) b = 1;
  return (int)(b > 148 * 4 ? 148 * 4 : b);  // 4 CTAs per SM: one CTA's tile loads run under the others' pixel loops
}

extern "C" size_t creste_wgrad_strided_workspace_bytes(int N, int P, int Q, int C, int K, int R, int S) {
  return (size_t)wgrad_strided_pb((long long)N * P * Q) * R * S * C * K * sizeof(double);
}

/* dw [R*S*C][K] of a strided dense conv with C == 4 input channels (the EfficientNet stem). */
extern "C" int creste_wgrad_strided(const float* x, const float* g, int N, int H, int W, int C, int K, int R, int S,
                                    int stride, int pad_t, int pad_l, int P, int Q, float* dw, void* ws,
                                    size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(x && g && dw && ws && N > 0 && H > 0 && W > 0 && K > 0 && R > 0 && S > 0 && stride > 0 && P > 0 &&
                       Q > 0, "creste_wgrad_strided: bad args");
  CRESTE_CHECK_ARG(C == 4 && R * S * K <= 1024, "creste_wgrad_strided: serves C == 4 and R*S*K <= 1024 only");
  CRESTE_CHECK_ARG(ws_bytes >= creste_wgrad_strided_workspace_bytes(N, P, Q, C, K, R, S), "creste_wgrad_strided: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int PB = wgrad_strided_pb((long long)N * P * Q);
  const int threads = (R * S * K + 31) / 32 * 32;
  const int tiles_q = ceil_div(Q, WG_TQ);
  const long long ntiles = (long long)N * P * tiles_q;
  const size_t smem = (size_t)R * ((WG_TQ - 1) * stride + S) * sizeof(float4) + (size_t)WG_TQ * K * sizeof(float);
  CRESTE_CHECK_ARG(smem <= 200 * 1024, "creste_wgrad_strided: tile does not fit shared memory");
  CRESTE_CUDA(cudaFuncSetAttribute(wgrad_strided_c4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_strided_c4_kernel<<<PB, threads, smem, st>>>(x, g, N, H, W, K, R, S, stride, pad_t, pad_l, P, Q, tiles_q, ntiles,
                                                    (double*)ws);
  int rc = launch_check("wgrad_strided_c4_kernel");
  if (rc) return rc;
  const int n = R * S * C * K;
  reduce_parts_kernel<float><<<ceil_div(n, 32), 256, 0, st>>>((const double*)ws, PB, n, 1, 1.0, dw);
  return launch_check("reduce_parts_kernel");
}

extern "C" int creste_bn_fwd_finalize(const double* stats2, const float* weight, const float* bias, int C, double M,
                                      double eps, float momentum, float* running_mean, float* running_var, float* ab,
                                      double* mean_inv, void* stream) {
  CRESTE_CHECK_ARG(stats2 && ab && mean_inv && C > 0 && M > 0 && (!running_mean == !running_var),
                   "creste_bn_fwd_finalize: bad args");
  bn_fwd_finalize_kernel<<<ceil_div(C, 256), 256, 0, (cudaStream_t)stream>>>(stats2, weight, bias, C, M, eps, momentum,
                                                                             running_mean, running_var, ab, mean_inv);
  return launch_check("bn_fwd_finalize_kernel");
}

extern "C" int creste_bn_bwd_finalize(const double* sums2, const float* ab, const double* mean_inv, int C, double M,
                                      float* out4, void* stream) {
  CRESTE_CHECK_ARG(sums2 && ab && mean_inv && out4 && C > 0 && M > 0, "creste_bn_bwd_finalize: bad args");
  bn_bwd_finalize_kernel<<<ceil_div(C, 256), 256, 0, (cudaStream_t)stream>>>(sums2, ab, mean_inv, C, M, out4);
  return launch_check("bn_bwd_finalize_kernel");
}

extern "C" int creste_wgrad_rows(const float* x, const float* g, int npix, int C, int K, float* dw, void* stream) {
  CRESTE_CHECK_ARG(x && g && dw && npix > 0 && npix <= 4096 && C > 0 && K > 0, "creste_wgrad_rows: bad args");
  wgrad_rows_kernel<<<ELT_GRID((long long)C * K), 256, 0, (cudaStream_t)stream>>>(x, g, npix, C, K, dw);
  return launch_check("wgrad_rows_kernel");
}

extern "C" int creste_ce_depth_bwd(const float* logits_nchw, const float* label_mm, int N, int D, long long HW,
                                   float depth_min, float depth_max, const float* scale_dev, float* dlogits,
                                   void* stream) {
  CRESTE_CHECK_ARG(logits_nchw && label_mm && scale_dev && dlogits && N > 0 && D > 0 && HW > 0,
                   "creste_ce_depth_bwd: bad args");
  const float bin_size = (float)(((double)depth_max - (double)depth_min) / (double)D);
  ce_depth_bwd_kernel<<<grid_cap2((long long)N * HW, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      logits_nchw, label_mm, N, D, HW, depth_min, bin_size, scale_dev, dlogits);
  return launch_check("ce_depth_bwd_kernel");
}

extern "C" int creste_masked_mse_bwd(const float* pred, const float* gt, long long n, const float* scale_dev,
                                     float* dpred, void* stream) {
  CRESTE_CHECK_ARG(pred && gt && scale_dev && dpred && n > 0, "creste_masked_mse_bwd: bad args");
  masked_mse_bwd_kernel<<<ELT_GRID(n), 256, 0, (cudaStream_t)stream>>>(pred, gt, n, scale_dev, dpred);
  return launch_check("masked_mse_bwd_kernel");
}
