// stage2.cu -- backward kernels of the stage-2 (train_ssc.py) surface: the bilinear BEV splat, the frustum
// un-projection and the soft-argmax depth, plus the two layout helpers the strided-convolution gradients use.
//
//   splat_bwd_voxel_kernel   per BEV cell: 1 / max(dens, min_w) and dD = Gd - [dens >= min_w] * <G, out> / dens
//   splat_bwd_point_kernel   per frustum point (one warp): gathers the 4 taps of G / max(dens, min_w) --
//                            d feats = sum_taps w * gF (a gather: no atomics), dw = <f, gF> + dD,
//                            d xy = sum_taps (d w / d r) * dw      (floor has no gradient)
//   frustum_bwd_kernel       d depth = <d xy, d xy / d depth> + d z * d z / d depth   (xy, z are affine in depth)
//   depth_expect_bwd_kernel  d logits_k = g * p_k * (val_k - E) / out_div  (softmax expectation)
//   dilate_kernel            zero insertion z[n, p*s, q*s, :] = g[n, p, q, :]  (data gradient of a strided conv =
//                            stride-1 conv of the dilated output gradient with the flipped weights)
//   phase_slice_kernel       x[:, a::s, b::s, :]  (weight gradient of a strided conv = stride-1 weight gradients of
//                            the s*s phase images)
//
// Formulas: oracle/splat_bwd_oracle.py (pinned to the reference's autograd of
// creste/models/blocks/splat_projection.py:262-354; golden tests/golden/splat_bwd.npz).
#include "common.cuh"

namespace creste {

// one warp per voxel; F % 4 == 0
__global__ void __launch_bounds__(256) splat_bwd_voxel_kernel(const float* __restrict__ g_bev,
                                                              const float* __restrict__ bev,
                                                              const float* __restrict__ dens,
                                                              const float* __restrict__ g_dens, long long NG, int F,
                                                              float min_w, float* __restrict__ inv,
                                                              float* __restrict__ dD) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (long long v = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); v < NG; v += (long long)gridDim.x * wpb) {
    const float d = dens[v];
    float s = 0.0f;
    if (d >= min_w) {        // below the clamp the normaliser is the constant min_w: no density gradient
      const float4* g4 = reinterpret_cast<const float4*>(g_bev + v * F);
      const float4* o4 = reinterpret_cast<const float4*>(bev + v * F);
      for (int c = lane; c < F / 4; c += 32) {
        const float4 a = __ldg(g4 + c), b = __ldg(o4 + c);
        s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); s = fmaf(a.w, b.w, s);
      }
      s = warp_sum(s);
    }
    if (lane == 0) {
      inv[v] = 1.0f / fmaxf(d, min_w);
      dD[v] = (g_dens ? g_dens[v] : 0.0f) - (d >= min_w ? s / d : 0.0f);
    }
  }
}

// one warp per point; F % 4 == 0, F <= 512
__global__ void __launch_bounds__(256) splat_bwd_point_kernel(const float* __restrict__ xy,
                                                              const float* __restrict__ feats,
                                                              const uint8_t* __restrict__ mask,
                                                              const float* __restrict__ g_bev,
                                                              const float* __restrict__ inv,
                                                              const float* __restrict__ dD, int N, int P, int F,
                                                              int H, int W, float* __restrict__ dfeats,
                                                              float* __restrict__ dxy) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const size_t G = (size_t)H * W;
  const int F4 = F / 4;
  for (long long pt = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pt < (long long)N * P;
       pt += (long long)gridDim.x * wpb) {
    const int n = (int)(pt / P);
    const float X = xy[pt * 2 + 0], Y = xy[pt * 2 + 1];
    const float fX = floorf(X), fY = floorf(Y);
    const long long X0 = (long long)fX, Y0 = (long long)fY;
    const float rX = __fsub_rn(X, (float)X0), rY = __fsub_rn(Y, (float)Y0);
    const bool m = mask ? (mask[pt] != 0) : true;
    const float4* f4 = reinterpret_cast<const float4*>(feats + (size_t)pt * F);
    float4 acc[4];                    // up to 4 channel groups per lane (F <= 512)
    float4 fv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int c = lane + 32 * j;
      fv[j] = (m && c < F4) ? __ldg(f4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float wX = __fadd_rn((float)(1 - dx), __fmul_rn((float)(2 * dx - 1), rX));
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const float wY = __fadd_rn((float)(1 - dy), __fmul_rn((float)(2 * dy - 1), rY));
        const long long X_ = X0 + dx, Y_ = Y0 + dy;
        if (!((0 <= X_) && (X_ < W) && (0 <= Y_) && (Y_ < H))) continue;      // warp-uniform
        const size_t id = (size_t)n * G + (size_t)(Y_ * W + X_);
        const float iv = __ldg(inv + id);
        const float w = __fmul_rn(wX, wY);
        const float4* g4 = reinterpret_cast<const float4*>(g_bev + id * F);
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = lane + 32 * j;
          if (c < F4) {
            float4 g = __ldg(g4 + c);
            g.x *= iv; g.y *= iv; g.z *= iv; g.w *= iv;
            acc[j].x = fmaf(w, g.x, acc[j].x); acc[j].y = fmaf(w, g.y, acc[j].y);
            acc[j].z = fmaf(w, g.z, acc[j].z); acc[j].w = fmaf(w, g.w, acc[j].w);
            dot = fmaf(fv[j].x, g.x, dot); dot = fmaf(fv[j].y, g.y, dot);
            dot = fmaf(fv[j].z, g.z, dot); dot = fmaf(fv[j].w, g.w, dot);
          }
        }
        dot = warp_sum(dot);
        const float dw = dot + __ldg(dD + id);
        gx = fmaf((float)(2 * dx - 1) * wY, dw, gx);
        gy = fmaf((float)(2 * dy - 1) * wX, dw, gy);
      }
    }
    float4* o4 = reinterpret_cast<float4*>(dfeats + (size_t)pt * F);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      if (c < F4) o4[c] = m ? acc[j] : make_float4(0.f, 0.f, 0.f, 0.f);     // masked points deposit density only
    }
    if (lane == 0) {
      dxy[pt * 2 + 0] = gx;
      dxy[pt * 2 + 1] = gy;
    }
  }
}

// xy0 = (-xmin - o1) / vx, xy1 = (-ymin - o0) / vy, z = o2 with o_r = M[r][0] u d + M[r][1] v d + M[r][2] d + M[r][3]
__global__ void frustum_bwd_kernel(const float* __restrict__ dxy, const float* __restrict__ dz,
                                   const float* __restrict__ p2p, int N, int Hs, int Ws, float vx, float vy,
                                   float* __restrict__ ddepth) {
  const int P = Hs * Ws;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * P) return;
  const int n = i / P, p = i - n * P;
  const int v = p / Ws, u = p - v * Ws;
  const float* M = p2p + n * 16;
  float d[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) d[r] = fmaf(M[r * 4 + 0], (float)u, fmaf(M[r * 4 + 1], (float)v, M[r * 4 + 2]));
  float g = 0.0f;
  if (dxy) g = -(dxy[(size_t)i * 2 + 0] * d[1]) / vx - (dxy[(size_t)i * 2 + 1] * d[0]) / vy;
  if (dz) g = fmaf(dz[i], d[2], g);
  ddepth[i] = g;
}

// one warp per pixel, D == 128: d logit_k = g * p_k * (val_k - E) / out_div
__global__ void __launch_bounds__(256) depth_expect_bwd_kernel(const float* __restrict__ logits,
                                                               const float* __restrict__ g_metric, int NP,
                                                               float dmin, float dmax, float out_div,
                                                               float* __restrict__ dlogits) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const float step = __fdiv_rn(__fsub_rn(dmax, dmin), 127.0f);
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < NP; p += gridDim.x * wpb) {
    const float4 l4 = __ldg(reinterpret_cast<const float4*>(logits + (size_t)p * 128) + lane);
    const float l[4] = {l4.x, l4.y, l4.z, l4.w};
    float m = warp_max(fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3])));
    float e[4], val[4], s = 0.0f, ev = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = lane * 4 + j;
      e[j] = expf(l[j] - m);
      val[j] = (k < 64) ? fmaf(step, (float)k, dmin) : dmax - step * (float)(127 - k);
      s += e[j];
      ev = fmaf(e[j], val[j], ev);
    }
    s = warp_sum(s);
    ev = warp_sum(ev);
    const float E = ev / s;
    const float g = g_metric[p] / (s * out_div);
    reinterpret_cast<float4*>(dlogits + (size_t)p * 128)[lane] =
        make_float4(g * e[0] * (val[0] - E), g * e[1] * (val[1] - E), g * e[2] * (val[2] - E), g * e[3] * (val[3] - E));
  }
}

// z [N,Hz,Wz,C] = 0 except z[n, p*s, q*s, :] = g[n,p,q,:]; C % 4 == 0
__global__ void __launch_bounds__(256) dilate_kernel(const float4* __restrict__ g, int N, int P, int Q, int C4,
                                                     int s, int Hz, int Wz, float4* __restrict__ z) {
  const long long total = (long long)N * Hz * Wz * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int x = (int)(t % Wz); t /= Wz;
    const int y = (int)(t % Hz);
    const int n = (int)(t / Hz);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y % s == 0 && x % s == 0 && y / s < P && x / s < Q)
      v = __ldg(g + (((size_t)n * P + y / s) * Q + x / s) * C4 + c);
    z[i] = v;
  }
}

// out [N,Ha,Wa,C] = x[n, a + s*i, b + s*j, :] (zero where the source index is outside the image); C % 4 == 0
__global__ void __launch_bounds__(256) phase_slice_kernel(const float4* __restrict__ x, int N, int H, int W, int C4,
                                                          int s, int a, int b, int Ha, int Wa,
                                                          float4* __restrict__ out) {
  const long long total = (long long)N * Ha * Wa * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long t = i / C4;
    const int xj = (int)(t % Wa); t /= Wa;
    const int yi = (int)(t % Ha);
    const int n = (int)(t / Ha);
    const int sy = a + s * yi, sx = b + s * xj;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = __ldg(x + (((size_t)n * H + sy) * W + sx) * C4 + c);
    out[i] = v;
  }
}

static inline int s2_grid(long long total, int threads, int cap) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

}  // namespace creste

using namespace creste;

extern "C" size_t creste_splat_bwd_workspace_bytes(int N, int H, int W) {
  return (size_t)2 * N * H * W * sizeof(float);
}

extern "C" int creste_splat_soft_bwd(const float* xy, const float* feats, const uint8_t* mask, const float* bev_nhwc,
                                     const float* dens, const float* g_bev_nhwc, const float* g_dens, int N, int P,
                                     int F, int H, int W, float min_weight, float* dfeats, float* dxy, void* ws,
                                     size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(xy && feats && bev_nhwc && dens && g_bev_nhwc && dfeats && dxy && ws,
                   "creste_splat_soft_bwd: null pointer");
  CRESTE_CHECK_ARG(N > 0 && P > 0 && H > 0 && W > 0 && F > 0 && F % 4 == 0 && F <= 512,
                   "creste_splat_soft_bwd: F must be a multiple of 4, <= 512");
  if (ws_bytes < creste_splat_bwd_workspace_bytes(N, H, W)) {
    set_error("creste_splat_soft_bwd: workspace too small");
    return CRESTE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long NG = (long long)N * H * W;
  float* inv = (float*)ws;
  float* dD = inv + NG;
  splat_bwd_voxel_kernel<<<s2_grid(NG, 8, 148 * 16), 256, 0, st>>>(g_bev_nhwc, bev_nhwc, dens, g_dens, NG, F,
                                                                   min_weight, inv, dD);
  int rc = launch_check("splat_bwd_voxel_kernel");
  if (rc) return rc;
  splat_bwd_point_kernel<<<s2_grid((long long)N * P, 8, 148 * 16), 256, 0, st>>>(xy, feats, mask, g_bev_nhwc, inv, dD,
                                                                                N, P, F, H, W, dfeats, dxy);
  return launch_check("splat_bwd_point_kernel");
}

extern "C" int creste_frustum_bwd(const float* dxy, const float* dz, const float* p2p, int N, int Hs, int Ws,
                                  const float* voxel, float* ddepth, void* stream) {
  CRESTE_CHECK_ARG((dxy || dz) && p2p && voxel && ddepth && N > 0 && Hs > 0 && Ws > 0, "creste_frustum_bwd: bad args");
  const int total = N * Hs * Ws;
  frustum_bwd_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(dxy, dz, p2p, N, Hs, Ws, voxel[0],
                                                                            voxel[1], ddepth);
  return launch_check("frustum_bwd_kernel");
}

extern "C" int creste_depth_expectation_bwd(const float* logits, const float* g_metric, int NP, int D,
                                            float depth_min_mm, float depth_max_mm, float out_div, float* dlogits,
                                            void* stream) {
  CRESTE_CHECK_ARG(logits && g_metric && dlogits && NP > 0, "creste_depth_expectation_bwd: bad args");
  CRESTE_CHECK_ARG(D == 128, "creste_depth_expectation_bwd: only D = 128 bins is implemented (got %d)", D);
  depth_expect_bwd_kernel<<<min(ceil_div(NP, 8), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      logits, g_metric, NP, depth_min_mm, depth_max_mm, out_div, dlogits);
  return launch_check("depth_expect_bwd_kernel");
}

extern "C" int creste_dilate(const float* g, int N, int P, int Q, int C, int stride, int Hz, int Wz, float* z,
                             void* stream) {
  CRESTE_CHECK_ARG(g && z && N > 0 && P > 0 && Q > 0 && C > 0 && C % 4 == 0 && stride >= 1 && Hz >= (P - 1) * stride + 1 &&
                       Wz >= (Q - 1) * stride + 1,
                   "creste_dilate: bad args");
  const long long total = (long long)N * Hz * Wz * (C / 4);
  dilate_kernel<<<s2_grid(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>((const float4*)g, N, P, Q, C / 4,
                                                                                stride, Hz, Wz, (float4*)z);
  return launch_check("dilate_kernel");
}

extern "C" int creste_phase_slice(const float* x, int N, int H, int W, int C, int stride, int a, int b, int Ha,
                                  int Wa, float* out, void* stream) {
  CRESTE_CHECK_ARG(x && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && stride >= 1 && Ha > 0 && Wa > 0,
                   "creste_phase_slice: bad args");
  const long long total = (long long)N * Ha * Wa * (C / 4);
  phase_slice_kernel<<<s2_grid(total, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, N, H, W, C / 4, stride, a, b, Ha, Wa, (float4*)out);
  return launch_check("phase_slice_kernel");
}
