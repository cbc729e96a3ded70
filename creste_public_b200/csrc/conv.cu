// conv.cu -- dispatcher of the conv family (creste_conv2d) + the EfficientNet trunk's
// memory-bound kernels (depthwise conv + BN + swish + SE partial sums, SE gate).
#include <stdlib.h>

#include "common.cuh"

namespace creste {
int conv_simt_launch(const creste_conv_desc* d, const float* x, const float* w, int ldw,
                     const float* scale, const float* shift, const float* gate,
                     const float* residual, float* out, unsigned* amax_out, cudaStream_t st);
int conv1x1_stream_launch(const creste_conv_desc* d, const float* x, const float* w, int ldw, const float* scale,
                          const float* shift, const float* gate, const float* residual, float* out,
                          unsigned* amax_out, cudaStream_t st);
struct TcSplitOut { void* hi; void* lo; float* scal; float bound_mul, bound_add; const float* in_amax; };   // conv_tc.cu
int conv_tc_launch(const creste_conv_desc* d, const float* x, const float* w_packed,
                   const float* scale, const float* shift, const float* gate, const float* residual,
                   float* out, const float* amax_in, unsigned* amax_out, void* ws, size_t ws_bytes,
                   cudaStream_t st, const TcSplitOut* so);
int conv_tc_presplit_launch(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                            const float* w_packed, const float* scale, const float* shift, const float* residual,
                            float* out, unsigned* amax_out, cudaStream_t st, const TcSplitOut* so);
size_t conv_tc_workspace_bytes(const creste_conv_desc* d);
bool conv_tc_supported(const creste_conv_desc* d);

// ---- depthwise conv + folded BN + swish + per-(n,c) spatial sum (for the SE block).
// NHWC: one thread per (pixel, 4-channel group); consecutive threads = consecutive channel
// groups, so every tap is a coalesced float4 load.  The SE sum is reduced in a FIXED order
// (per-thread sequential -> shared memory across the block's pixel lanes -> one partial row per
// block; creste_se_gate adds the rows in order): no atomics, bit-reproducible run to run, which
// matters because the encoder->depth->splat chain amplifies 1e-7 perturbations.
template <int R>
__global__ void __launch_bounds__(256) dwconv_kernel(const float* __restrict__ x,
                                                     const float* __restrict__ w,
                                                     const float* __restrict__ scale,
                                                     const float* __restrict__ shift, int N, int H,
                                                     int W, int C, int stride, int pad_t, int pad_l,
                                                     int P, int Q, float* __restrict__ out,
                                                     float* __restrict__ chan_part) {
  // grid: x = pixel tiles within one image, y = image, z = chunks of 256 channel groups;
  // chan_part [N, gridDim.x, C]
  __shared__ float4 s_part[256];
  const int C4 = C / 4;
  const int n = blockIdx.y;
  const int PQ = P * Q;
  const int pix_per_block = gridDim.x > 0 ? ceil_div(PQ, (int)gridDim.x) : PQ;
  const int pix0 = blockIdx.x * pix_per_block;
  const int pix1 = min(pix0 + pix_per_block, PQ);
  // thread -> channel group cg = t % C4s ; pixel lane = t / C4s   (C4s = min(C4, 256))
  {
    const int cg0 = blockIdx.z * 256;
    const int cgs = min(C4 - cg0, 256);
    const int lanes = 256 / cgs;  // pixels processed concurrently by the block
    const int cg = cg0 + (threadIdx.x % cgs);
    const int pl = threadIdx.x / cgs;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pl < lanes) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cg);
      const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
      float4 wk[R * R];
#pragma unroll
      for (int i = 0; i < R * R; ++i) wk[i] = __ldg(reinterpret_cast<const float4*>(w + (size_t)i * C) + cg);
      for (int pix = pix0 + pl; pix < pix1; pix += lanes) {
        const int oy = pix / Q, ox = pix - oy * Q;
        const int iy0 = oy * stride - pad_t, ix0 = ox * stride - pad_l;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int iy = iy0 + r;
          if (iy < 0 || iy >= H) continue;
#pragma unroll
          for (int s = 0; s < R; ++s) {
            const int ix = ix0 + s;
            if (ix < 0 || ix >= W) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)n * H + iy) * W + ix) * C) + cg);
            const float4 k = wk[r * R + s];
            acc.x = fmaf(v.x, k.x, acc.x); acc.y = fmaf(v.y, k.y, acc.y);
            acc.z = fmaf(v.z, k.z, acc.z); acc.w = fmaf(v.w, k.w, acc.w);
          }
        }
        float4 o;
        o.x = fmaf(acc.x, sc.x, sh.x); o.y = fmaf(acc.y, sc.y, sh.y);
        o.z = fmaf(acc.z, sc.z, sh.z); o.w = fmaf(acc.w, sc.w, sh.w);
        o.x = o.x / (1.0f + expf(-o.x)); o.y = o.y / (1.0f + expf(-o.y));
        o.z = o.z / (1.0f + expf(-o.z)); o.w = o.w / (1.0f + expf(-o.w));
        reinterpret_cast<float4*>(out + ((size_t)n * PQ + pix) * C)[cg] = o;
        sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
      }
    }
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (pl == 0 && threadIdx.x < cgs) {
      float4 tot = s_part[threadIdx.x];
      for (int l = 1; l < lanes; ++l) {
        const float4 o = s_part[l * cgs + threadIdx.x];
        tot.x += o.x; tot.y += o.y; tot.z += o.z; tot.w += o.w;
      }
      reinterpret_cast<float4*>(chan_part + ((size_t)n * gridDim.x + blockIdx.x) * C)[cg] = tot;
    }
  }
}

// ---- x-blocked variant: a thread produces XB adjacent outputs of one row for its 4 channels and
// loads each input column once per filter row ((XB-1)*STRIDE + R float4 loads for XB*R FMAs-by-4
// instead of XB*R): the previous kernel issued 25 loads per 5x5 output and measured ~6x its HBM floor.
// Same FMA order per output (taps in (r,s) raster order; out-of-image taps add an exact 0), same
// fixed-order SE partial sums (per-thread sequential -> lanes in order -> one row per block).
template <int R, int STRIDE, int XB>
__global__ void __launch_bounds__(256) dwconv_xb_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int N, int H, int W,
                                                        int C, int pad_t, int pad_l, int P, int Q,
                                                        float* __restrict__ out, float* __restrict__ chan_part,
                                                        unsigned* __restrict__ amax_out) {
  constexpr int NC = (XB - 1) * STRIDE + R;
  __shared__ float4 s_part[256];
  const int C4 = C / 4;
  const int n = blockIdx.y;
  const int Qb = (Q + XB - 1) / XB;
  const int units = P * Qb;
  const int per_block = ceil_div(units, (int)gridDim.x);
  const int u0 = blockIdx.x * per_block;
  const int u1 = min(u0 + per_block, units);
  const int cg0 = blockIdx.z * 256;
  const int cgs = min(C4 - cg0, 256);
  const int lanes = 256 / cgs;
  const int cg = cg0 + (threadIdx.x % cgs);
  const int pl = threadIdx.x / cgs;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  float amx = 0.0f;
  if (pl < lanes) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cg);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
    const float4* wv = reinterpret_cast<const float4*>(w) + cg;
    const float4* xn = reinterpret_cast<const float4*>(x + (size_t)n * H * W * C) + cg;
    for (int u = u0 + pl; u < u1; u += lanes) {
      const int oy = u / Qb, ox0 = (u - oy * Qb) * XB;
      const int iy0 = oy * STRIDE - pad_t, ix0 = ox0 * STRIDE - pad_l;
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
      for (int b = 0; b < XB; ++b) {
        if (ox0 + b < Q) {
          float4 o;
          o.x = fmaf(acc[b].x, sc.x, sh.x); o.y = fmaf(acc[b].y, sc.y, sh.y);
          o.z = fmaf(acc[b].z, sc.z, sh.z); o.w = fmaf(acc[b].w, sc.w, sh.w);
          o.x = o.x / (1.0f + expf(-o.x)); o.y = o.y / (1.0f + expf(-o.y));
          o.z = o.z / (1.0f + expf(-o.z)); o.w = o.w / (1.0f + expf(-o.w));
          reinterpret_cast<float4*>(out + (((size_t)n * P + oy) * Q + ox0 + b) * C)[cg] = o;
          sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
          amx = fmaxf(fmaxf(amx, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        }
      }
    }
  }
  // max|out| travels with the tensor: the project conv derives its 3xFP16 operand scale from it (the SE gate is a
  // sigmoid, so max|out * gate| <= max|out|) instead of making an amax pass over out * gate
  if (amax_out) {
    amx = warp_max(amx);
    if ((threadIdx.x & 31) == 0 && amx > 0.0f) atomicMax(amax_out, __float_as_uint(amx));
  }
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (pl == 0 && threadIdx.x < cgs) {
    float4 tot = s_part[threadIdx.x];
    for (int l = 1; l < lanes; ++l) {
      const float4 o = s_part[l * cgs + threadIdx.x];
      tot.x += o.x; tot.y += o.y; tot.z += o.z; tot.w += o.w;
    }
    reinterpret_cast<float4*>(chan_part + ((size_t)n * gridDim.x + blockIdx.x) * C)[cg] = tot;
  }
}

// ---- tiled variant (round 2b): the x-blocked kernel above reads every tap straight from global memory with per-column
// bounds checks and re-derives its indices per unit -- ncu: 272 instructions per float4 output, issue-bound at 0.25-0.35
// of the HBM rate.  Here a CTA stages the input tile of TH x TW outputs x 32 channels (with its halo, zero-filled outside
// the image: exactly the static TF-SAME padding) in shared memory with coalesced loads, and the taps come from there
// without checks.  Same FMA order per output (taps in (r,s) raster order from acc = 0, out-of-image taps add an exact
// 0), same epilogue; the SE partial sums are again reduced in a fixed order (thread: its units in order, outputs left
// to right; CTA: the 32 unit lanes in order; one row of chan_part per tile), so results are reproducible run to run and
// independent of the batch size -- but the rows are tiles now, so the squeeze-excite MEAN is summed in a different
// order than with the x-blocked kernel (last-bit differences; both are within the parity bars).
template <int R, int STRIDE>
__global__ void __launch_bounds__(256) dwconv_tile_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ scale, const float* __restrict__ shift,
                                                          int H, int W, int C, int pad_t, int pad_l, int P, int Q,
                                                          int tiles_x, float* __restrict__ out,
                                                          float* __restrict__ chan_part, unsigned* __restrict__ amax_out) {
  constexpr int TH = 8, TW = STRIDE == 1 ? 32 : 16, XB = 4, CG = 8;
  constexpr int IH = (TH - 1) * STRIDE + R, IW = (TW - 1) * STRIDE + R;
  constexpr int NC = (XB - 1) * STRIDE + R;
  constexpr int UNITS = TH * (TW / XB);           // 64 (stride 1) / 32 (stride 2) units of XB outputs
  extern __shared__ __align__(16) float4 dw_smem[];
  float4* tile = dw_smem;                          // [IH][IW][CG]
  float4* wsm = dw_smem + IH * IW * CG;            // [R*R][CG]
  __shared__ float4 s_part[256];
  const int C4 = C / 4;
  const int n = blockIdx.y;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int oy0 = ty * TH, ox0 = tx * TW;
  const int cg0 = blockIdx.z * CG;
  const int cgs = min(CG, C4 - cg0);
  const int tid = threadIdx.x;
  const float4* xn = reinterpret_cast<const float4*>(x + (size_t)n * H * W * C);
  const int iy0 = oy0 * STRIDE - pad_t, ix0 = ox0 * STRIDE - pad_l;
  for (int i = tid; i < IH * IW * CG; i += 256) {
    const int g = i % CG, p = i / CG;
    const int ix = ix0 + p % IW, iy = iy0 + p / IW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < cgs && iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xn + ((size_t)iy * W + ix) * C4 + cg0 + g);
    tile[i] = v;
  }
  for (int i = tid; i < R * R * CG; i += 256) {
    const int g = i % CG, t = i / CG;
    wsm[i] = g < cgs ? __ldg(reinterpret_cast<const float4*>(w) + (size_t)t * C4 + cg0 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int g = tid % CG, ul = tid / CG;           // 32 unit lanes
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  float amx = 0.0f;
  if (g < cgs) {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + cg0 + g);
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + cg0 + g);
#pragma unroll 1
    for (int u = ul; u < UNITS; u += 32) {
      const int ly = u / (TW / XB), lx = (u - ly * (TW / XB)) * XB;
      const int oy = oy0 + ly;
      if (oy >= P || ox0 + lx >= Q) continue;
      float4 acc[XB];
#pragma unroll
      for (int b = 0; b < XB; ++b) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* t0 = tile + ((size_t)(ly * STRIDE) * IW + lx * STRIDE) * CG + g;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float4 col[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) col[c] = t0[((size_t)r * IW + c) * CG];
#pragma unroll
        for (int s_ = 0; s_ < R; ++s_) {
          const float4 k = wsm[(r * R + s_) * CG + g];
#pragma unroll
          for (int b = 0; b < XB; ++b) {
            const float4 v = col[b * STRIDE + s_];
            acc[b].x = fmaf(v.x, k.x, acc[b].x); acc[b].y = fmaf(v.y, k.y, acc[b].y);
            acc[b].z = fmaf(v.z, k.z, acc[b].z); acc[b].w = fmaf(v.w, k.w, acc[b].w);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < XB; ++b) {
        if (ox0 + lx + b < Q) {
          float4 o;
          o.x = fmaf(acc[b].x, sc.x, sh.x); o.y = fmaf(acc[b].y, sc.y, sh.y);
          o.z = fmaf(acc[b].z, sc.z, sh.z); o.w = fmaf(acc[b].w, sc.w, sh.w);
          o.x = o.x / (1.0f + expf(-o.x)); o.y = o.y / (1.0f + expf(-o.y));
          o.z = o.z / (1.0f + expf(-o.z)); o.w = o.w / (1.0f + expf(-o.w));
          reinterpret_cast<float4*>(out + (((size_t)n * P + oy) * Q + ox0 + lx + b) * C)[cg0 + g] = o;
          sum.x += o.x; sum.y += o.y; sum.z += o.z; sum.w += o.w;
          amx = fmaxf(fmaxf(amx, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        }
      }
    }
  }
  if (amax_out) {
    amx = warp_max(amx);
    if ((tid & 31) == 0 && amx > 0.0f) atomicMax(amax_out, __float_as_uint(amx));
  }
  s_part[tid] = sum;
  __syncthreads();
  if (tid < cgs) {
    float4 tot = s_part[tid];
    for (int l = 1; l < 32; ++l) {
      const float4 o = s_part[l * CG + tid];
      tot.x += o.x; tot.y += o.y; tot.z += o.z; tot.w += o.w;
    }
    reinterpret_cast<float4*>(chan_part + ((size_t)n * gridDim.x + blockIdx.x) * C)[cg0 + tid] = tot;
  }
}

// ---- SE gate: one block per image
__global__ void __launch_bounds__(1024) se_gate_kernel(const float* __restrict__ chan_part, int nparts,
                                                      float inv_hw, int C, int Csq,
                                                      const float* __restrict__ w_red,
                                                      const float* __restrict__ b_red,
                                                      const float* __restrict__ w_exp,
                                                      const float* __restrict__ b_exp,
                                                      float* __restrict__ gate) {
  extern __shared__ float sm[];  // mean[C], sq[Csq], lane sums [8][C]
  float* mean = sm;
  float* sq = sm + C;
  float* lsum = sm + C + Csq;
  const int n = blockIdx.x;
  // fixed summation order, independent of the batch size: 8 row lanes per channel add rows lane, lane + 8,
  // ... in order, then the lane sums are added in lane order (a single thread per channel walking all the
  // partial rows made this kernel 7 % of the forward in the round-1 launch list)
  for (int idx = threadIdx.x; idx < 8 * C; idx += blockDim.x) {
    const int l = idx / C, c = idx - l * C;
    const float* src = chan_part + (size_t)n * nparts * C + c;
    float tot = 0.0f;
    for (int j = l; j < nparts; j += 8) tot += __ldg(src + (size_t)j * C);
    lsum[idx] = tot;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float tot = lsum[c];
#pragma unroll
    for (int l = 1; l < 8; ++l) tot += lsum[l * C + c];
    mean[c] = tot * inv_hw;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < Csq; j += nw) {
    float a = 0.0f;
    for (int c = lane; c < C; c += 32) a = fmaf(w_red[(size_t)j * C + c], mean[c], a);
    a = warp_sum(a);
    if (lane == 0) {
      a += b_red[j];
      sq[j] = a / (1.0f + expf(-a));  // swish
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = b_exp[c];
    for (int j = 0; j < Csq; ++j) a = fmaf(w_exp[(size_t)c * Csq + j], sq[j], a);
    gate[(size_t)n * C + c] = 1.0f / (1.0f + expf(-a));
  }
}

// ---- SE gate across a thread-block cluster: 8 CTAs per image, CTA r owns channels [r*Cs, (r+1)*Cs).  The
// one-CTA form above is latency-bound on its first phase (one SM pulling up to 592 x 1152 partial sums out of L2:
// 37-48 us per call, 10 % of the B = 1 frame); here that phase and the excite phase run 8-wide, the per-channel
// means are exchanged through distributed shared memory, and every CTA evaluates the (tiny) squeeze layer itself.
// Same summation order per channel as se_gate_kernel (rows lane, lane + 8, ... then the lanes in order), so the
// gate is bit-identical.
constexpr int SE_CL = 8;
__global__ void __cluster_dims__(SE_CL, 1, 1) __launch_bounds__(512)
se_gate_cluster_kernel(const float* __restrict__ chan_part, int nparts, float inv_hw, int C, int Csq,
                       const float* __restrict__ w_red, const float* __restrict__ b_red,
                       const float* __restrict__ w_exp, const float* __restrict__ b_exp, float* __restrict__ gate) {
  extern __shared__ float sm[];  // mean[C] (all channels, filled by every CTA of the cluster), sq[Csq], lsum[8][Cs]
  const int Cs = (C + SE_CL - 1) / SE_CL;
  float* mean = sm;
  float* sq = sm + C;
  float* lsum = sm + C + Csq;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int n = blockIdx.x / SE_CL;
  const int c0 = (int)rank * Cs, cn = max(0, min(Cs, C - c0));
  for (int idx = threadIdx.x; idx < 8 * cn; idx += blockDim.x) {
    const int l = idx / cn, c = idx - l * cn;
    const float* src = chan_part + (size_t)n * nparts * C + c0 + c;
    float tot = 0.0f;
    for (int j = l; j < nparts; j += 8) tot += __ldg(src + (size_t)j * C);
    lsum[l * Cs + c] = tot;
  }
  __syncthreads();
  // own means -> every CTA's `mean` array (distributed shared memory stores)
  const unsigned mean_addr = (unsigned)__cvta_generic_to_shared(mean);
  for (int c = threadIdx.x; c < cn; c += blockDim.x) {
    float tot = lsum[c];
#pragma unroll
    for (int l = 1; l < 8; ++l) tot += lsum[l * Cs + c];
    const float m = tot * inv_hw;
#pragma unroll
    for (unsigned r = 0; r < SE_CL; ++r) {
      unsigned ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(mean_addr + (unsigned)(c0 + c) * 4u), "r"(r));
      asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(m) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < Csq; j += nw) {
    float a = 0.0f;
    for (int c = lane; c < C; c += 32) a = fmaf(w_red[(size_t)j * C + c], mean[c], a);
    a = warp_sum(a);
    if (lane == 0) {
      a += b_red[j];
      sq[j] = a / (1.0f + expf(-a));  // swish
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cn; c += blockDim.x) {
    float a = b_exp[c0 + c];
    for (int j = 0; j < Csq; ++j) a = fmaf(w_exp[(size_t)(c0 + c) * Csq + j], sq[j], a);
    gate[(size_t)n * C + c0 + c] = 1.0f / (1.0f + expf(-a));
  }
  // no CTA may exit while a peer can still store into its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace creste

using namespace creste;

extern "C" size_t creste_conv2d_workspace_bytes(const creste_conv_desc* d) {
  if (!d || d->precision == 0) return 0;
  return conv_tc_workspace_bytes(d);
}

extern "C" int creste_conv2d(const creste_conv_desc* d, const float* x, const float* w_packed,
                             const float* scale, const float* shift, const float* gate,
                             const float* residual, float* out, void* ws, size_t ws_bytes,
                             void* stream) {
  return creste_conv2d_ex(d, x, w_packed, scale, shift, gate, residual, out, nullptr, nullptr, ws, ws_bytes, stream);
}

static int conv2d_impl(const creste_conv_desc* d, const float* x, const float* w_packed, const float* scale,
                       const float* shift, const float* gate, const float* residual, float* out, const float* amax_in,
                       float* amax_out, const TcSplitOut* so, void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(d && x && w_packed && (out || (so && so->hi)), "creste_conv2d: null pointer");
  CRESTE_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->K > 0 && d->R > 0 && d->S > 0 &&
                       d->stride > 0 && d->P > 0 && d->Q > 0,
                   "creste_conv2d: bad shape");
  CRESTE_CHECK_ARG(d->C % 4 == 0, "creste_conv2d: C must be a multiple of 4 (got %d)", d->C);
  CRESTE_CHECK_ARG((d->P - 1) * d->stride - d->pad_t + d->R - 1 >= 0, "creste_conv2d: bad padding");
  cudaStream_t st = (cudaStream_t)stream;
  if (amax_out) CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), st));
  if (d->precision == 0) {
    CRESTE_CHECK_ARG(!(so && so->hi), "creste_conv2d: the split output needs a tensor-core precision mode");
    const int ldw = (d->K + 3) / 4 * 4;
    // HBM-bound 1x1 convs with short reductions: the streaming kernel (same arithmetic, bit-identical results)
    const int rc = conv1x1_stream_launch(d, x, w_packed, ldw, scale, shift, gate, residual, out, (unsigned*)amax_out, st);
    if (rc != 0) return rc == 1 ? 0 : rc;
    return conv_simt_launch(d, x, w_packed, ldw, scale, shift, gate, residual, out, (unsigned*)amax_out, st);
  }
  if (!conv_tc_supported(d)) {
    set_error("creste_conv2d: precision mode %d is not available for this shape "
              "(C=%d K=%d R=%d stride=%d); use precision 0", d->precision, d->C, d->K, d->R, d->stride);
    return CRESTE_ERR_ARG;
  }
  return conv_tc_launch(d, x, w_packed, scale, shift, gate, residual, out, amax_in, (unsigned*)amax_out, ws, ws_bytes, st, so);
}

extern "C" int creste_conv2d_ex(const creste_conv_desc* d, const float* x, const float* w_packed,
                                const float* scale, const float* shift, const float* gate,
                                const float* residual, float* out, const float* amax_in, float* amax_out,
                                void* ws, size_t ws_bytes, void* stream) {
  return conv2d_impl(d, x, w_packed, scale, shift, gate, residual, out, amax_in, amax_out, nullptr, ws, ws_bytes, stream);
}

extern "C" int creste_conv2d_split_out(const creste_conv_desc* d, const float* x, const float* w_packed,
                                       const float* scale, const float* shift, const float* gate,
                                       const float* residual, float* out, const float* amax_in, float* amax_out,
                                       void* out_hi, void* out_lo, float* out_scal, float bound_mul, float bound_add,
                                       void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(out_hi && out_scal, "creste_conv2d_split_out: null pointer");
  const TcSplitOut so = {out_hi, out_lo, out_scal, bound_mul, bound_add, amax_in};
  return conv2d_impl(d, x, w_packed, scale, shift, gate, residual, out, amax_in, amax_out, &so, ws, ws_bytes, stream);
}

static int conv2d_presplit_impl(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                                const float* w_packed, const float* scale, const float* shift, const float* residual,
                                float* out, float* amax_out, const TcSplitOut* so, void* stream) {
  CRESTE_CHECK_ARG(d && x_hi && x_scal && w_packed && (out || (so && so->hi)), "creste_conv2d_presplit: null pointer");
  CRESTE_CHECK_ARG(d->precision == 4 || d->precision == 5, "creste_conv2d_presplit: precision must be 4 (3xFP16) or 5 (fp16)");
  CRESTE_CHECK_ARG(d->precision == 5 || x_lo, "creste_conv2d_presplit: the 3xFP16 mode needs the lo halves");
  if (!conv_tc_supported(d)) {
    set_error("creste_conv2d_presplit: shape not served by the tensor-core kernel (C=%d K=%d R=%d stride=%d)", d->C,
              d->K, d->R, d->stride);
    return CRESTE_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (amax_out) CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), st));
  return conv_tc_presplit_launch(d, x_hi, x_lo, x_scal, w_packed, scale, shift, residual, out, (unsigned*)amax_out, st, so);
}

extern "C" int creste_conv2d_presplit(const creste_conv_desc* d, const void* x_hi, const void* x_lo,
                                      const float* x_scal, const float* w_packed, const float* scale,
                                      const float* shift, const float* residual, float* out, float* amax_out,
                                      void* stream) {
  return conv2d_presplit_impl(d, x_hi, x_lo, x_scal, w_packed, scale, shift, residual, out, amax_out, nullptr, stream);
}

extern "C" int creste_conv2d_presplit_split_out(const creste_conv_desc* d, const void* x_hi, const void* x_lo,
                                                const float* x_scal, const float* w_packed, const float* scale,
                                                const float* shift, const float* residual, float* out, float* amax_out,
                                                void* out_hi, void* out_lo, float* out_scal, float bound_mul,
                                                float bound_add, const float* x_amax, void* stream) {
  CRESTE_CHECK_ARG(out_hi && out_scal, "creste_conv2d_presplit_split_out: null pointer");
  const TcSplitOut so = {out_hi, out_lo, out_scal, bound_mul, bound_add, x_amax};
  return conv2d_presplit_impl(d, x_hi, x_lo, x_scal, w_packed, scale, shift, residual, out, amax_out, &so, stream);
}

extern "C" int creste_dwconv_num_parts(int N, int P, int Q) {
  // independent of N on purpose: the per-image summation order (hence the result bits) must not
  // depend on how many frames share the launch (frames are the data-parallel sharding unit).
  // >= 2 CTAs per SM worth of tiles even for the 16x30 layers, <= 64 pixels per tile.
  (void)N;
  const int PQ = P * Q;
  int tiles = ceil_div(PQ, 64);
  if (tiles < 296) tiles = PQ < 296 ? PQ : 296;
  return tiles > 592 ? 592 : tiles;
}

extern "C" int creste_dwconv_tile_parts(int P, int Q, int stride) {
  // chan_part rows of the tiled depthwise kernel: one per 8 x 32 (stride 1) / 8 x 16 (stride 2) output tile
  return ceil_div(P, 8) * ceil_div(Q, stride == 1 ? 32 : 16);
}

extern "C" int creste_dwconv_parts(int N, int P, int Q, int R, int stride) {
  // the chan_part row count of the kernel that is faster for this layer (measured, tools/kernel_bench.py): the tiled
  // kernel wins on the low-resolution stages (<= 32 x 60 outputs) and on the 5 x 5 stride-1 layers up to 64 x 120; the
  // x-blocked one on the large early layers (its L1-served taps beat the tile fill there)
  const long long pq = (long long)P * Q;
  const bool tiled = !getenv("CRESTE_NO_DWTILE") && (stride == 1 || stride == 2) &&
                     (pq <= 2048 || (R == 5 && stride == 1 && pq <= 7680) || getenv("CRESTE_DWTILE_ALL"));
  const int t = creste_dwconv_tile_parts(P, Q, stride), o = creste_dwconv_num_parts(N, P, Q);
  return tiled ? t : o;
}

extern "C" int creste_dwconv_bn_swish(const float* x, const float* w, const float* scale,
                                      const float* shift, int N, int H, int W, int C, int R,
                                      int stride, int pad_t, int pad_l, int P, int Q, float* out,
                                      float* chan_part, int nparts, void* stream) {
  return creste_dwconv_bn_swish_ex(x, w, scale, shift, N, H, W, C, R, stride, pad_t, pad_l, P, Q, out, chan_part, nparts,
                                   nullptr, stream);
}

extern "C" int creste_dwconv_bn_swish_ex(const float* x, const float* w, const float* scale,
                                         const float* shift, int N, int H, int W, int C, int R,
                                         int stride, int pad_t, int pad_l, int P, int Q, float* out,
                                         float* chan_part, int nparts, float* amax_out, void* stream) {
  CRESTE_CHECK_ARG(x && w && scale && shift && out && chan_part, "creste_dwconv_bn_swish: null pointer");
  CRESTE_CHECK_ARG(C % 4 == 0 && (R == 3 || R == 5), "creste_dwconv_bn_swish: C%%4==0, R in {3,5}");
  CRESTE_CHECK_ARG(nparts == creste_dwconv_num_parts(N, P, Q) || nparts == creste_dwconv_tile_parts(P, Q, stride),
                   "creste_dwconv_bn_swish: nparts");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* am = (unsigned*)amax_out;
  if (am) CRESTE_CUDA(cudaMemsetAsync(am, 0, sizeof(float), st));
  // tiled kernel: selected by the caller through the chan_part row count it allocated (creste_dwconv_tile_parts)
  if ((stride == 1 || stride == 2) && nparts == creste_dwconv_tile_parts(P, Q, stride) &&
      nparts != creste_dwconv_num_parts(N, P, Q)) {
    const int TW = stride == 1 ? 32 : 16;
    const int tiles_x = ceil_div(Q, TW);
    const int IH = 7 * stride + R, IW = (TW - 1) * stride + R;
    const size_t smem = ((size_t)IH * IW * 8 + (size_t)R * R * 8) * sizeof(float4);
    dim3 tgrid(nparts, N, ceil_div(C / 4, 8));
#define CRESTE_DWT(RR, SS)                                                                                            \
    do {                                                                                                              \
      CRESTE_CUDA(cudaFuncSetAttribute(dwconv_tile_kernel<RR, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      dwconv_tile_kernel<RR, SS><<<tgrid, 256, smem, st>>>(x, w, scale, shift, H, W, C, pad_t, pad_l, P, Q, tiles_x, out, \
                                                          chan_part, am);                                            \
    } while (0)
    if (stride == 1 && R == 3) CRESTE_DWT(3, 1);
    else if (stride == 1) CRESTE_DWT(5, 1);
    else if (R == 3) CRESTE_DWT(3, 2);
    else CRESTE_DWT(5, 2);
#undef CRESTE_DWT
    return launch_check("dwconv_tile_kernel");
  }
  dim3 grid(nparts, N, ceil_div(C / 4, 256));
  if (stride == 1 && R == 3)
    dwconv_xb_kernel<3, 1, 4><<<grid, 256, 0, st>>>(x, w, scale, shift, N, H, W, C, pad_t, pad_l, P, Q, out, chan_part, am);
  else if (stride == 1 && R == 5)
    dwconv_xb_kernel<5, 1, 4><<<grid, 256, 0, st>>>(x, w, scale, shift, N, H, W, C, pad_t, pad_l, P, Q, out, chan_part, am);
  else if (stride == 2 && R == 3)
    dwconv_xb_kernel<3, 2, 4><<<grid, 256, 0, st>>>(x, w, scale, shift, N, H, W, C, pad_t, pad_l, P, Q, out, chan_part, am);
  else if (stride == 2 && R == 5)
    dwconv_xb_kernel<5, 2, 4><<<grid, 256, 0, st>>>(x, w, scale, shift, N, H, W, C, pad_t, pad_l, P, Q, out, chan_part, am);
  else {
    CRESTE_CHECK_ARG(!am, "creste_dwconv_bn_swish_ex: amax_out needs stride 1 or 2");
    if (R == 3)
      dwconv_kernel<3><<<grid, 256, 0, st>>>(x, w, scale, shift, N, H, W, C, stride, pad_t, pad_l, P, Q, out, chan_part);
    else
      dwconv_kernel<5><<<grid, 256, 0, st>>>(x, w, scale, shift, N, H, W, C, stride, pad_t, pad_l, P, Q, out, chan_part);
  }
  return launch_check("dwconv_kernel");
}

extern "C" int creste_se_gate(const float* chan_part, int nparts, float inv_hw, int N, int C, int Csq,
                              const float* w_red, const float* b_red, const float* w_exp,
                              const float* b_exp, float* gate, void* stream) {
  CRESTE_CHECK_ARG(chan_part && w_red && b_red && w_exp && b_exp && gate && nparts > 0, "creste_se_gate: null pointer");
  if (!getenv("CRESTE_SE_ONE_CTA")) {
    const int Cs = (C + SE_CL - 1) / SE_CL;
    const size_t smem = (size_t)(C + Csq + 8 * Cs) * sizeof(float);
    se_gate_cluster_kernel<<<N * SE_CL, 512, smem, (cudaStream_t)stream>>>(chan_part, nparts, inv_hw, C, Csq, w_red,
                                                                          b_red, w_exp, b_exp, gate);
    return launch_check("se_gate_cluster_kernel");
  }
  const size_t smem = (size_t)(9 * C + Csq) * sizeof(float);
  se_gate_kernel<<<N, 1024, smem, (cudaStream_t)stream>>>(chan_part, nparts, inv_hw, C, Csq, w_red, b_red, w_exp,
                                                        b_exp, gate);
  return launch_check("se_gate_kernel");
}
