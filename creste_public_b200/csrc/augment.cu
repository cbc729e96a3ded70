// augment.cu -- on-device input pipeline: the per-frame warps the reference's CPU
// dataloader runs through kornia / cv2 before a frame reaches the model.
//
//   creste_affine_warp    kornia.geometry.transform.warp_affine as the reference calls it
//                         (creste/utils/utils.py:6-38 `warp`: a ones-channel is appended, warped with the map, and
//                         thresholded at 0.99 into the validity mask; used by RotateAndTranslate.transform_map,
//                         creste/utils/train_utils.py:213-232, for the BEV labels and the FOV-mask pose warp,
//                         creste/datasets/codapefree_dataloader.py:691-709)
//   creste_depth_augment  DepthAugmentation.__call__ (creste/utils/train_utils.py:110-181) in ONE pass:
//                         LiDAR dropout (mask = u > p), camera-LiDAR miscalibration (bilinear affine warp of the
//                         dropped map, zeros outside) and additive Gaussian noise, given the random draws
//   creste_traverse_to_bev  CodaPEFreeDataset._load_traverse (codapefree_dataloader.py:579-615): LiDAR-frame SE(3)
//                         poses of the expert trajectory -> clamped 3x3 BEV grid poses
//
// warp_affine = F.affine_grid(theta) + F.grid_sample (zeros padding): the kernels take theta, the 2x3 map from
// NORMALISED output coordinates to NORMALISED input coordinates, which the host derives from the pixel-space matrix
// exactly as kornia does (creste/utils/train_utils.py of the mirror, `affine_theta`).  Arithmetic follows ATen's
// grid sampler: base grid (2j + 1) / W - 1 (align_corners = false) or linspace(-1, 1, W) (true); the source
// coordinate ((g + 1) * W - 1) / 2 or (g + 1) / 2 * (W - 1); bilinear weights from floor(); nearest = nearbyint
// (round half to even).  HBM-bound: one read of the taps (L2-resident neighbourhood) and one write per output element.
#include <math_constants.h>

#include "common.cuh"

namespace creste {

__device__ __forceinline__ float aug_base(int j, int n, int align) {
  // ATen linspace_from_neg_one: align_corners: -1 + 2 j / (n - 1); else the same scaled by (n - 1) / n
  if (n <= 1) return 0.0f;
  // torch.linspace(-1, 1, n): start + step * i in the first half, end - step * (n - 1 - i) in the second
  const float step = 2.0f / (float)(n - 1);
  const float v = j < n / 2 ? -1.0f + step * (float)j : 1.0f - step * (float)(n - 1 - j);
  return align ? v : v * (float)(n - 1) / (float)n;
}
__device__ __forceinline__ float aug_unnorm(float g, int n, int align) {
  return align ? ((g + 1.0f) / 2.0f) * (float)(n - 1) : ((g + 1.0f) * (float)n - 1.0f) / 2.0f;
}

struct AugTaps {
  int x0, y0, x1, y1;
  float w00, w01, w10, w11;   // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
  bool nearest_ok; int xn, yn;
};

__device__ __forceinline__ AugTaps aug_taps(const float* th, int ox, int oy, int H, int W, int Ho, int Wo, int align) {
  const float xb = aug_base(ox, Wo, align), yb = aug_base(oy, Ho, align);
  // affine_grid: base [x, y, 1] @ theta^T (a 3-term dot product; ATen evaluates it as a batched matmul)
  const float gx = fmaf(th[2], 1.0f, fmaf(th[1], yb, th[0] * xb));
  const float gy = fmaf(th[5], 1.0f, fmaf(th[4], yb, th[3] * xb));
  const float ix = aug_unnorm(gx, W, align), iy = aug_unnorm(gy, H, align);
  AugTaps t;
  const float fx = floorf(ix), fy = floorf(iy);
  t.x0 = (int)fx; t.y0 = (int)fy; t.x1 = t.x0 + 1; t.y1 = t.y0 + 1;
  const float x1f = fx + 1.0f, y1f = fy + 1.0f;          // ATen: nw = (ix_se - ix) * (iy_se - iy), ...
  t.w00 = (x1f - ix) * (y1f - iy); t.w01 = (ix - fx) * (y1f - iy);
  t.w10 = (x1f - ix) * (iy - fy);  t.w11 = (ix - fx) * (iy - fy);
  const float rx = nearbyintf(ix), ry = nearbyintf(iy);
  t.xn = (int)rx; t.yn = (int)ry;
  t.nearest_ok = rx >= 0.0f && rx <= (float)(W - 1) && ry >= 0.0f && ry <= (float)(H - 1);
  return t;
}

__device__ __forceinline__ bool aug_in(int x, int y, int H, int W) { return x >= 0 && x < W && y >= 0 && y < H; }

// in [B][C][H][W] -> out [B][C][Ho][Wo], mask [B][Ho][Wo] (warped ones-channel > 0.99); theta [B][6]
__global__ void __launch_bounds__(256) affine_warp_kernel(const float* __restrict__ in, int B, int C, int H, int W,
                                                          const float* __restrict__ theta, int Ho, int Wo, int nearest,
                                                          int align, float* __restrict__ out,
                                                          unsigned char* __restrict__ mask) {
  const long long total = (long long)B * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const long long r = i / Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    const AugTaps t = aug_taps(theta + b * 6, ox, oy, H, W, Ho, Wo, align);
    const float* src = in + (size_t)b * C * H * W;
    float* dst = out + (size_t)b * C * Ho * Wo + (size_t)oy * Wo + ox;
    if (nearest) {
      for (int c = 0; c < C; ++c)
        dst[(size_t)c * Ho * Wo] = t.nearest_ok ? __ldg(src + ((size_t)c * H + t.yn) * W + t.xn) : 0.0f;
      if (mask) mask[i] = t.nearest_ok ? 1 : 0;
    } else {
      const bool i00 = aug_in(t.x0, t.y0, H, W), i01 = aug_in(t.x1, t.y0, H, W);
      const bool i10 = aug_in(t.x0, t.y1, H, W), i11 = aug_in(t.x1, t.y1, H, W);
      for (int c = 0; c < C; ++c) {
        const float* p = src + (size_t)c * H * W;
        float v = 0.0f;                                  // ATen order: nw, ne, sw, se accumulated into the output
        if (i00) v += __ldg(p + (size_t)t.y0 * W + t.x0) * t.w00;
        if (i01) v += __ldg(p + (size_t)t.y0 * W + t.x1) * t.w01;
        if (i10) v += __ldg(p + (size_t)t.y1 * W + t.x0) * t.w10;
        if (i11) v += __ldg(p + (size_t)t.y1 * W + t.x1) * t.w11;
        dst[(size_t)c * Ho * Wo] = v;
      }
      if (mask) {
        float m = 0.0f;
        if (i00) m += t.w00;
        if (i01) m += t.w01;
        if (i10) m += t.w10;
        if (i11) m += t.w11;
        mask[i] = m > 0.99f ? 1 : 0;
      }
    }
  }
}

// out = warp_bilinear(depth * (u > p_drop), theta; align_corners = true, zeros) + g * noise_std     (all [H][W])
__global__ void __launch_bounds__(256) depth_augment_kernel(const float* __restrict__ depth, const float* __restrict__ u,
                                                            const float* __restrict__ g, int H, int W, float p_drop,
                                                            const float* __restrict__ theta, float noise_std,
                                                            float* __restrict__ out) {
  const long long total = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W), oy = (int)(i / W);
    const AugTaps t = aug_taps(theta, ox, oy, H, W, H, W, 1);
    auto tap = [&](int x, int y) -> float {
      const size_t o = (size_t)y * W + x;
      // depth_map * mask with mask = rand > p (a bool promoted to 0 / 1)
      return __ldg(depth + o) * ((__ldg(u + o) > p_drop) ? 1.0f : 0.0f);
    };
    float v = 0.0f;
    if (aug_in(t.x0, t.y0, H, W)) v += tap(t.x0, t.y0) * t.w00;
    if (aug_in(t.x1, t.y0, H, W)) v += tap(t.x1, t.y0) * t.w01;
    if (aug_in(t.x0, t.y1, H, W)) v += tap(t.x0, t.y1) * t.w10;
    if (aug_in(t.x1, t.y1, H, W)) v += tap(t.x1, t.y1) * t.w11;
    out[i] = v + __ldg(g + i) * noise_std;
  }
}

// poses [T][4][4] (LiDAR frame, relative to the first) -> grid poses [T][3][3]:
//   P = eye(3); P[:2,:2] = R[:2,:2]; P[:2,2] = t[:2] / voxel;  G = T_lidar_to_bev @ P with
//   T_lidar_to_bev = [[-1, 0, W // 2], [0, -1, H // 2], [0, 0, 1]];  G[:2,2] clamped to [0, (H, W)]
// (every product is by -1, 0 or 1 and every sum has at most two non-zero terms: exact in any order)
__global__ void traverse_to_bev_kernel(const float* __restrict__ poses, int T, float vx, float vy, int bevH, int bevW,
                                       float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* P = poses + (size_t)t * 16;
  float* G = out + (size_t)t * 9;
  const float cx = (float)(bevW / 2), cy = (float)(bevH / 2);
  const float px = P[3] / vx, py = P[7] / vy;
  G[0] = -P[0]; G[1] = -P[1];
  G[3] = -P[4]; G[4] = -P[5];
  G[2] = fminf(fmaxf(-px + cx, 0.0f), (float)bevH);
  G[5] = fminf(fmaxf(-py + cy, 0.0f), (float)bevW);
  G[6] = 0.0f; G[7] = 0.0f; G[8] = 1.0f;
}

static int aug_grid(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace creste

using namespace creste;

extern "C" int creste_affine_warp(const float* in, int B, int C, int H, int W, const float* theta, int Ho, int Wo,
                                  int nearest, int align_corners, float* out, unsigned char* mask, void* stream) {
  CRESTE_CHECK_ARG(in && theta && out && B > 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "creste_affine_warp: bad args");
  affine_warp_kernel<<<aug_grid((long long)B * Ho * Wo), 256, 0, (cudaStream_t)stream>>>(in, B, C, H, W, theta, Ho, Wo,
                                                                                        nearest, align_corners, out, mask);
  return launch_check("affine_warp_kernel");
}

extern "C" int creste_depth_augment(const float* depth, const float* u, const float* g, int H, int W, float p_drop,
                                    const float* theta, float noise_std, float* out, void* stream) {
  CRESTE_CHECK_ARG(depth && u && g && theta && out && H > 0 && W > 0, "creste_depth_augment: bad args");
  depth_augment_kernel<<<aug_grid((long long)H * W), 256, 0, (cudaStream_t)stream>>>(depth, u, g, H, W, p_drop, theta,
                                                                                    noise_std, out);
  return launch_check("depth_augment_kernel");
}

extern "C" int creste_traverse_to_bev(const float* poses, int T, float voxel_x, float voxel_y, int bev_h, int bev_w,
                                      float* out, void* stream) {
  CRESTE_CHECK_ARG(poses && out && T > 0 && voxel_x > 0 && voxel_y > 0, "creste_traverse_to_bev: bad args");
  traverse_to_bev_kernel<<<(T + 127) / 128, 128, 0, (cudaStream_t)stream>>>(poses, T, voxel_x, voxel_y, bev_h, bev_w, out);
  return launch_check("traverse_to_bev_kernel");
}
