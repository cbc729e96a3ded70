// splat.cu -- camera frustum -> BEV bilinear splat (K6-K9).
//
// Replaces Camera2World.forward (reference creste/models/blocks/splat_projection.py:19-51), the
// bounds mask (:169), _points_to_voxels (:175-189), the z-MLP + concat (:152-158) and
// splat_soft (:262-354).  The reference issues a meshgrid + H2D upload + bmm, ~40 elementwise
// launches and 8 scatter_add_ launches per frame.
//
//   frustum_kernel   one thread per frustum point: K=4 in-order FMA chain (the CPU bmm order),
//                    bounds mask, lidar2map (one inexact add per coordinate), TRUE division by the
//                    voxel size -- the integer voxel indices derived from xy are bit-exact.
//   zmlp_concat      one warp per point: copies the 256 image features (float4, coalesced NHWC)
//                    and evaluates the 1->64->32 MLP with the weights in shared memory.
//   splat_kernel     one warp per point, channels across lanes (float4 per lane, 128 B-coalesced):
//                    4 taps x F channels of `red.global.add.v4.f32` into an NHWC accumulator;
//                    lane 0 adds the tap weight to the density plane.
//   normalize_kernel accumulator / clamp(density, min_weight) -> NHWC (decoder input) and/or NCHW
//                    (the reference's `bev_features` layout), one pass.
//
// HBM roofline: 37.6 MB/frame algorithmic at F=96, P=30720, 256x256.
#include "common.cuh"

namespace creste {

__global__ void frustum_kernel(const float* __restrict__ depth, const float* __restrict__ p2p,
                               int N, int Hs, int Ws, float xmin, float ymin, float zmin,
                               float xmax, float ymax, float zmax, float vx, float vy,
                               float* __restrict__ xy, float* __restrict__ zout,
                               uint8_t* __restrict__ mask) {
  const int P = Hs * Ws;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * P) return;
  const int n = i / P, p = i - n * P;
  const int v = p / Ws, u = p - v * Ws;
  const float* M = p2p + n * 16;
  const float d = depth[i];
  // campts = [u*d, v*d, 1*d, 1]   (splat_projection.py:37-45)
  const float c0 = __fmul_rn((float)u, d), c1 = __fmul_rn((float)v, d), c2 = d;
  float o[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(M[r * 4 + 0], c0);
    acc = __fmaf_rn(M[r * 4 + 1], c1, acc);
    acc = __fmaf_rn(M[r * 4 + 2], c2, acc);
    acc = __fmaf_rn(M[r * 4 + 3], 1.0f, acc);
    o[r] = acc;
  }
  const bool ok = (o[0] < xmax) && (o[0] >= xmin) && (o[1] < ymax) && (o[1] >= ymin) &&
                  (o[2] < zmax) && (o[2] >= zmin);
  if (mask) mask[i] = ok ? 1 : 0;
  if (zout) zout[i] = o[2];
  // lidar2map = [[0,-1,0,-xmin],[-1,0,0,-ymin],...] (splat_projection.py:81-88), then / voxel
  const float xm = __fadd_rn(-xmin, -o[1]);
  const float ym = __fadd_rn(-ymin, -o[0]);
  xy[(size_t)i * 2 + 0] = __fdiv_rn(xm, vx);
  xy[(size_t)i * 2 + 1] = __fdiv_rn(ym, vy);
}

// Camera2World.forward as a stand-alone op (the fused path never materialises xyz): xyz [N,3,Hs,Ws]
__global__ void camera_to_world_kernel(const float* __restrict__ depth, const float* __restrict__ p2p, int N,
                                       int Hs, int Ws, float* __restrict__ xyz) {
  const int P = Hs * Ws;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * P) return;
  const int n = i / P, p = i - n * P;
  const int v = p / Ws, u = p - v * Ws;
  const float* M = p2p + n * 16;
  const float d = depth[i];
  const float c0 = __fmul_rn((float)u, d), c1 = __fmul_rn((float)v, d), c2 = d;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(M[r * 4 + 0], c0);
    acc = __fmaf_rn(M[r * 4 + 1], c1, acc);
    acc = __fmaf_rn(M[r * 4 + 2], c2, acc);
    acc = __fmaf_rn(M[r * 4 + 3], 1.0f, acc);
    xyz[((size_t)n * 3 + r) * P + p] = acc;
  }
}

// Camera2MapMulti._points_to_voxels: rows 0..1 of (lidar2map @ [x,y,z,1]) / voxel_size[:2]
__global__ void points_to_voxels_kernel(const float* __restrict__ pts, long long NP, float l00, float l01,
                                        float l02, float l03, float l10, float l11, float l12, float l13,
                                        float vx, float vy, float* __restrict__ xy) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NP) return;
  const float x = pts[i * 3 + 0], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
  float a = __fmul_rn(l00, x);
  a = __fmaf_rn(l01, y, a); a = __fmaf_rn(l02, z, a); a = __fmaf_rn(l03, 1.0f, a);
  float b = __fmul_rn(l10, x);
  b = __fmaf_rn(l11, y, b); b = __fmaf_rn(l12, z, b); b = __fmaf_rn(l13, 1.0f, b);
  xy[i * 2 + 0] = __fdiv_rn(a, vx);
  xy[i * 2 + 1] = __fdiv_rn(b, vy);
}

// one warp per point; C % 4 == 0
__global__ void __launch_bounds__(256) zmlp_concat_kernel(const float* __restrict__ feats,
                                                          const float* __restrict__ z, int NP, int C,
                                                          const float* __restrict__ w1,
                                                          const float* __restrict__ b1,
                                                          const float* __restrict__ w2,
                                                          const float* __restrict__ b2,
                                                          float* __restrict__ out, unsigned* __restrict__ amax_out) {
  __shared__ float s_w1[64], s_b1[64], s_b2[32];
  __shared__ float s_w2[32 * 65];  // padded: lane j reads row j
  for (int i = threadIdx.x; i < 64; i += blockDim.x) { s_w1[i] = w1[i]; s_b1[i] = b1[i]; }
  for (int i = threadIdx.x; i < 32; i += blockDim.x) s_b2[i] = b2[i];
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) s_w2[(i / 64) * 65 + (i % 64)] = w2[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  float amx = 0.0f;
  for (int pt = blockIdx.x * warps_per_block + (threadIdx.x >> 5); pt < NP;
       pt += gridDim.x * warps_per_block) {
    const float4* src = reinterpret_cast<const float4*>(feats + (size_t)pt * C);
    float4* dst = reinterpret_cast<float4*>(out + (size_t)pt * (C + 32));
    for (int c = lane; c < C / 4; c += 32) {
      const float4 v = __ldg(src + c);
      dst[c] = v;
      amx = fmaxf(fmaxf(amx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    const float zz = z[pt];
    // hidden layer: lane computes h[lane], h[lane+32]  (Linear(1,64) + ReLU)
    const float h0 = fmaxf(__fmaf_rn(s_w1[lane], zz, s_b1[lane]), 0.0f);
    const float h1 = fmaxf(__fmaf_rn(s_w1[lane + 32], zz, s_b1[lane + 32]), 0.0f);
    // output j = lane: sum_k w2[j][k] h[k] + b2[j]  (Linear(64,32) + ReLU), k in order
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = __fmaf_rn(s_w2[lane * 65 + k], __shfl_sync(0xffffffffu, h0, k), acc);
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = __fmaf_rn(s_w2[lane * 65 + 32 + k], __shfl_sync(0xffffffffu, h1, k), acc);
    const float zf = fmaxf(acc + s_b2[lane], 0.0f);
    out[(size_t)pt * (C + 32) + C + lane] = zf;
    amx = fmaxf(amx, zf);
  }
  // max|out| travels with the tensor: the fusion conv derives its 3xFP16 operand scale from it (no amax pass)
  if (amax_out) {
    amx = warp_max(amx);
    if (lane == 0 && amx > 0.0f) atomicMax(amax_out, __float_as_uint(amx));
  }
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// one warp per point; F % 4 == 0
__global__ void __launch_bounds__(256) splat_kernel(const float* __restrict__ xy,
                                                    const float* __restrict__ feats,
                                                    const uint8_t* __restrict__ mask, int N, int P,
                                                    int F, int H, int W, float* __restrict__ acc,
                                                    float* __restrict__ dens,
                                                    long long* __restrict__ idx_out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const size_t G = (size_t)H * W;
  for (long long pt = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pt < (long long)N * P;
       pt += (long long)gridDim.x * wpb) {
    const int n = (int)(pt / P);
    const float X = xy[pt * 2 + 0], Y = xy[pt * 2 + 1];
    // XY = floor().long(); rXY = xy - XY   (splat_projection.py:293-294)
    const float fX = floorf(X), fY = floorf(Y);
    const long long X0 = (long long)fX, Y0 = (long long)fY;
    const float rX = __fsub_rn(X, (float)X0), rY = __fsub_rn(Y, (float)Y0);
    const bool m = mask ? (mask[pt] != 0) : true;
    float* accn = acc + (size_t)n * G * F;
    float* densn = dens + (size_t)n * G;
    const float4* fsrc = reinterpret_cast<const float4*>(feats + (size_t)pt * F);
    int t = 0;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float wX = __fadd_rn((float)(1 - dx), __fmul_rn((float)(2 * dx - 1), rX));
#pragma unroll
      for (int dy = 0; dy < 2; ++dy, ++t) {
        const float wY = __fadd_rn((float)(1 - dy), __fmul_rn((float)(2 * dy - 1), rY));
        const float w = __fmul_rn(wX, wY);
        const long long X_ = X0 + dx, Y_ = Y0 + dy;
        const bool valid = (0 <= X_) && (X_ < W) && (0 <= Y_) && (Y_ < H);
        const long long id = Y_ * W + X_;
        if (idx_out && lane == 0) idx_out[pt * 4 + t] = valid ? id : -1;
        if (!valid) continue;
        if (lane == 0) atomicAdd(densn + id, w);  // density also for masked points (:219, :331)
        if (m && accn) {
          for (int c = lane; c < F / 4; c += 32) {
            float4 f = __ldg(fsrc + c);
            f.x = __fmul_rn(w, f.x); f.y = __fmul_rn(w, f.y);
            f.z = __fmul_rn(w, f.z); f.w = __fmul_rn(w, f.w);
            red_add_v4(accn + (size_t)id * F + c * 4, f);
          }
        }
      }
    }
  }
}

// acc NHWC [N,G,F] / clamp(dens) -> nhwc and/or nchw.  Tile transpose through shared memory so
// both the NHWC read and the NCHW write are coalesced.
__global__ void __launch_bounds__(256) splat_normalize_kernel(const float* __restrict__ acc,
                                                              const float* __restrict__ dens, int N,
                                                              int G, int F, float min_w,
                                                              float* __restrict__ nhwc,
                                                              float* __restrict__ nchw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int g0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int g = g0 + j, f = f0 + tx;
    float v = 0.0f;
    if (g < G && f < F) {
      const float d = dens[(size_t)n * G + g];
      v = __fdiv_rn(acc[((size_t)n * G + g) * F + f], fmaxf(d, min_w));
      if (nhwc) nhwc[((size_t)n * G + g) * F + f] = v;
    }
    tile[j][tx] = v;
  }
  __syncthreads();
  if (nchw) {
    for (int j = ty; j < 32; j += 8) {
      const int f = f0 + j, g = g0 + tx;
      if (g < G && f < F) nchw[((size_t)n * F + f) * G + g] = tile[tx][j];
    }
  }
}

}  // namespace creste

using namespace creste;

extern "C" int creste_frustum_to_bev(const float* depth, const float* p2p, int N, int Hs, int Ws,
                                     const float* range, const float* voxel, float* xy, float* z,
                                     uint8_t* mask, void* stream) {
  CRESTE_CHECK_ARG(depth && p2p && range && voxel && xy, "creste_frustum_to_bev: null pointer");
  CRESTE_CHECK_ARG(N > 0 && Hs > 0 && Ws > 0, "creste_frustum_to_bev: bad shape");
  const int total = N * Hs * Ws;
  frustum_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      depth, p2p, N, Hs, Ws, range[0], range[1], range[2], range[3], range[4], range[5], voxel[0],
      voxel[1], xy, z, mask);
  return launch_check("frustum_kernel");
}

extern "C" int creste_camera_to_world(const float* depth, const float* p2p, int N, int Hs, int Ws, float* xyz,
                                      void* stream) {
  CRESTE_CHECK_ARG(depth && p2p && xyz, "creste_camera_to_world: null pointer");
  CRESTE_CHECK_ARG(N > 0 && Hs > 0 && Ws > 0, "creste_camera_to_world: bad shape");
  const int total = N * Hs * Ws;
  camera_to_world_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(depth, p2p, N, Hs, Ws, xyz);
  return launch_check("camera_to_world_kernel");
}

extern "C" int creste_points_to_voxels(const float* pts, long long NP, const float* lidar2map, const float* voxel,
                                       float* xy, void* stream) {
  CRESTE_CHECK_ARG(pts && lidar2map && voxel && xy && NP > 0, "creste_points_to_voxels: null pointer");
  const float* L = lidar2map;   // HOST 4x4 row-major
  points_to_voxels_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pts, NP, L[0], L[1], L[2], L[3], L[4], L[5], L[6], L[7], voxel[0], voxel[1], xy);
  return launch_check("points_to_voxels_kernel");
}

extern "C" int creste_zmlp_concat(const float* feats, const float* z, int NP, int C, const float* w1,
                                  const float* b1, const float* w2, const float* b2, float* out,
                                  void* stream) {
  return creste_zmlp_concat_ex(feats, z, NP, C, w1, b1, w2, b2, out, nullptr, stream);
}

extern "C" int creste_zmlp_concat_ex(const float* feats, const float* z, int NP, int C, const float* w1,
                                     const float* b1, const float* w2, const float* b2, float* out,
                                     float* amax_out, void* stream) {
  CRESTE_CHECK_ARG(feats && z && w1 && b1 && w2 && b2 && out, "creste_zmlp_concat: null pointer");
  CRESTE_CHECK_ARG(NP > 0 && C > 0 && C % 4 == 0, "creste_zmlp_concat: C must be a multiple of 4");
  if (amax_out) CRESTE_CUDA(cudaMemsetAsync(amax_out, 0, sizeof(float), (cudaStream_t)stream));
  const int blocks = min(ceil_div(NP, 8), 148 * 8);
  zmlp_concat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(feats, z, NP, C, w1, b1, w2, b2, out,
                                                             (unsigned*)amax_out);
  return launch_check("zmlp_concat_kernel");
}

extern "C" size_t creste_splat_workspace_bytes(int N, int H, int W, int F) {
  return (size_t)N * H * W * F * sizeof(float);
}

extern "C" int creste_splat_soft(const float* xy, const float* feats, const uint8_t* mask, int N,
                                 int P, int F, int H, int W, float min_weight, float* bev_nhwc,
                                 float* bev_nchw, float* dens, int64_t* idx_out, void* ws,
                                 size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(xy && dens, "creste_splat_soft: null pointer");
  CRESTE_CHECK_ARG(N > 0 && P > 0 && H > 0 && W > 0, "creste_splat_soft: bad shape");
  const bool want_feats = (bev_nhwc || bev_nchw);
  if (want_feats) {
    CRESTE_CHECK_ARG(feats && F > 0 && F % 4 == 0, "creste_splat_soft: F must be a multiple of 4");
    if (!ws || ws_bytes < creste_splat_workspace_bytes(N, H, W, F)) {
      set_error("creste_splat_soft: workspace too small");
      return CRESTE_ERR_WORKSPACE;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t G = (size_t)H * W;
  CRESTE_CUDA(cudaMemsetAsync(dens, 0, (size_t)N * G * sizeof(float), st));
  if (want_feats) CRESTE_CUDA(cudaMemsetAsync(ws, 0, (size_t)N * G * F * sizeof(float), st));
  const long long total = (long long)N * P;
  const int blocks = (int)((total + 7) / 8 < 148 * 16 ? (total + 7) / 8 : 148 * 16);
  splat_kernel<<<blocks, 256, 0, st>>>(xy, feats, mask, N, P, F, H, W,
                                       want_feats ? (float*)ws : nullptr, dens,
                                       (long long*)idx_out);
  int rc = launch_check("splat_kernel");
  if (rc) return rc;
  if (want_feats) {
    dim3 grid(ceil_div((int)G, 32), ceil_div(F, 32), N);
    splat_normalize_kernel<<<grid, 256, 0, st>>>((const float*)ws, dens, N, (int)G, F, min_weight,
                                                 bev_nhwc, bev_nchw);
    rc = launch_check("splat_normalize_kernel");
  }
  return rc;
}
