// losses2.cu -- the stage-2 (train_ssc.py) losses as fused kernels: values and gradients.
//
//   smooth_l1        masked Smooth-L1 mean (SmoothL1Depth on the soft-argmax depth, loss_utils.py:530-573; SmoothL1 on
//                    the elevation head, :576-604): {sum, count} in double + the gradient pass
//   ce_weighted      class-weighted cross-entropy over the FOV-masked BEV cells (CrossEntropy, :379-474):
//                    {sum w*nll, sum w, #correct among gt != 0, #(gt != 0)} + the gradient pass (NCHW logits)
//   supcon           multi-positive contrastive loss (SupPixelConLoss :203-286 -> MultiPosConLoss,
//                    creste/models/losses/supcon_loss.py:56-115) over L2-normalised pixel embeddings: one pass of
//                    online log-sum-exp over the N x Na similarity matrix (never materialised), and the two
//                    gradient passes (rows: local embeddings; columns: the all-gathered embeddings)
//
// Reductions: per-thread partials -> warp shuffle -> one double atomicAdd per block (the loss scalars are sums of
// ~1e5 terms; the summation order changes the last bits only, as in PyTorch's own CUDA reductions).
#include "common.cuh"

namespace creste {

static inline int l2_grid(long long total, int threads, int cap) {
  long long b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > cap ? cap : b);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------- smooth L1
// valid = mask[i] (if given) && isfinite(gt[i]);  d = pred[i] - gt[i] * gt_scale
__global__ void __launch_bounds__(256) smooth_l1_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                        const uint8_t* __restrict__ mask, long long n, float gt_scale,
                                                        float beta, double* __restrict__ acc) {
  double s = 0.0, c = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = __ldg(gt + i);
    if ((mask && !mask[i]) || !isfinite(g)) continue;
    const float d = fabsf(__ldg(pred + i) - g * gt_scale);
    s += (double)(d < beta ? 0.5f * d * d / beta : d - 0.5f * beta);
    c += 1.0;
  }
  s = warp_sum_d(s);
  c = warp_sum_d(c);
  if ((threadIdx.x & 31) == 0) { atomicAdd(acc, s); atomicAdd(acc + 1, c); }
}

__global__ void __launch_bounds__(256) smooth_l1_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            const uint8_t* __restrict__ mask, long long n, float gt_scale,
                                                            float beta, const float* __restrict__ scale_dev,
                                                            float* __restrict__ dpred) {
  const float sc = __ldg(scale_dev);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = __ldg(gt + i);
    float o = 0.0f;
    if (!(mask && !mask[i]) && isfinite(g)) {
      const float d = __ldg(pred + i) - g * gt_scale;
      o = sc * (fabsf(d) < beta ? d / beta : (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)));
    }
    dpred[i] = o;
  }
}

// --------------------------------------------------------------------------- weighted cross-entropy
// logits NCHW [B,C,HW]; labels [B,HW] int64; mask [B,HW] uint8 (NULL = all); weights [C] (NULL = 1)
__global__ void __launch_bounds__(256) ce_weighted_kernel(const float* __restrict__ logits,
                                                          const long long* __restrict__ labels,
                                                          const uint8_t* __restrict__ mask,
                                                          const float* __restrict__ weights, int B, int C, long long HW,
                                                          long long ignore_index, double* __restrict__ acc) {
  double nll = 0.0, wsum = 0.0, ok = 0.0, cnt = 0.0;
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;
    const long long b = i / HW, p = i - b * HW;
    const long long y = labels[i];
    const float* src = logits + (size_t)b * C * HW + p;
    float m = -INFINITY;
    int am = 0;
    for (int k = 0; k < C; ++k) {
      const float v = __ldg(src + (size_t)k * HW);
      if (v > m) { m = v; am = k; }
    }
    if (y != 0) { cnt += 1.0; ok += (am == (int)y) ? 1.0 : 0.0; }      // mIoU meta: 0 is the "unlabeled" class
    if (y == ignore_index || y < 0 || y >= C) continue;
    float ssum = 0.f;
    for (int k = 0; k < C; ++k) ssum += expf(__ldg(src + (size_t)k * HW) - m);
    const float w = weights ? __ldg(weights + y) : 1.0f;
    nll += (double)(w * (logf(ssum) + m - __ldg(src + (size_t)y * HW)));
    wsum += (double)w;
  }
  nll = warp_sum_d(nll); wsum = warp_sum_d(wsum); ok = warp_sum_d(ok); cnt = warp_sum_d(cnt);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc, nll); atomicAdd(acc + 1, wsum); atomicAdd(acc + 2, ok); atomicAdd(acc + 3, cnt);
  }
}

__global__ void __launch_bounds__(256) ce_weighted_bwd_kernel(const float* __restrict__ logits,
                                                              const long long* __restrict__ labels,
                                                              const uint8_t* __restrict__ mask,
                                                              const float* __restrict__ weights, int B, int C,
                                                              long long HW, long long ignore_index,
                                                              const float* __restrict__ scale_dev,
                                                              float* __restrict__ dlogits) {
  const float sc = __ldg(scale_dev);
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* src = logits + (size_t)b * C * HW + p;
    float* dst = dlogits + (size_t)b * C * HW + p;
    const long long y = labels[i];
    const bool live = !(mask && !mask[i]) && y != ignore_index && y >= 0 && y < C;
    if (!live) {
      for (int k = 0; k < C; ++k) dst[(size_t)k * HW] = 0.0f;
      continue;
    }
    float m = -INFINITY;
    for (int k = 0; k < C; ++k) m = fmaxf(m, __ldg(src + (size_t)k * HW));
    float ssum = 0.f;
    for (int k = 0; k < C; ++k) ssum += expf(__ldg(src + (size_t)k * HW) - m);
    const float w = (weights ? __ldg(weights + y) : 1.0f) * sc;
    for (int k = 0; k < C; ++k) {
      const float pk = expf(__ldg(src + (size_t)k * HW) - m) / ssum;
      dst[(size_t)k * HW] = w * (pk - (k == (int)y ? 1.0f : 0.0f));
    }
  }
}

// ------------------------------------------------------------------- multi-positive contrastive loss
// f [N,D] local embeddings, a [Na,D] all (gathered) embeddings, both L2-NORMALISED; la / lb int64 labels;
// row i of the local block is column (i + self_off) of the gathered block (masked out of both the softmax
// and the positives).  logits_ij = <f_i, a_j> / T.
constexpr int SC_TI = 64;       // rows per CTA
constexpr int SC_TJ = 64;       // columns per smem tile
constexpr int SC_DMAX = 128;

// stats[i] = {row max m_i, sum_j exp(l_ij - m_i), #positives c_i, sum over positives of l_ij}
template <int D4>
__global__ void __launch_bounds__(256) supcon_rows_kernel(const float* __restrict__ f, const float* __restrict__ a,
                                                          const long long* __restrict__ lf,
                                                          const long long* __restrict__ la, int N, int Na,
                                                          int self_off, float inv_t, float4* __restrict__ stats) {
  __shared__ float4 s_a[SC_TJ][D4 + 1];
  __shared__ long long s_l[SC_TJ];
  const int i0 = blockIdx.x * SC_TI;
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;       // 64 rows x 4 column lanes
  const int i = i0 + r;
  float4 fi[D4];
#pragma unroll
  for (int d = 0; d < D4; ++d)
    fi[d] = (i < N) ? __ldg(reinterpret_cast<const float4*>(f + (size_t)i * D4 * 4) + d) : make_float4(0.f, 0.f, 0.f, 0.f);
  const long long li = (i < N) ? lf[i] : -1;
  float m = -INFINITY, s = 0.f, cpos = 0.f, spos = 0.f;
  for (int j0 = 0; j0 < Na; j0 += SC_TJ) {
    __syncthreads();
    for (int t = threadIdx.x; t < SC_TJ * D4; t += blockDim.x) {
      const int jj = t / D4, d = t - jj * D4;
      s_a[jj][d] = (j0 + jj < Na) ? __ldg(reinterpret_cast<const float4*>(a + (size_t)(j0 + jj) * D4 * 4) + d)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < SC_TJ) s_l[threadIdx.x] = (j0 + threadIdx.x < Na) ? la[j0 + threadIdx.x] : -2;
    __syncthreads();
    if (i < N) {
      for (int jj = q; jj < SC_TJ; jj += 4) {
        const int j = j0 + jj;
        if (j >= Na || j == i + self_off) continue;
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < D4; ++d) {
          const float4 av = s_a[jj][d];
          dot = fmaf(fi[d].x, av.x, dot); dot = fmaf(fi[d].y, av.y, dot);
          dot = fmaf(fi[d].z, av.z, dot); dot = fmaf(fi[d].w, av.w, dot);
        }
        const float l = dot * inv_t;
        if (l > m) { s = s * expf(m - l) + 1.0f; m = l; } else { s += expf(l - m); }
        if (s_l[jj] == li) { cpos += 1.0f; spos += l; }
      }
    }
  }
  // merge the 4 column lanes of a row (adjacent lanes of one warp)
#pragma unroll
  for (int o = 1; o < 4; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float c2 = __shfl_xor_sync(0xffffffffu, cpos, o), p2 = __shfl_xor_sync(0xffffffffu, spos, o);
    const float mm = fmaxf(m, m2);
    s = (mm == -INFINITY) ? 0.f : s * expf(m - mm) + s2 * expf(m2 - mm);
    m = mm; cpos += c2; spos += p2;
  }
  if (i < N && q == 0) stats[i] = make_float4(m, s, cpos, spos);
}

// loss = mean_i w_i * ( (c_i > 0) * (m_i + log s_i) - spos_i / max(c_i, 1) )
__global__ void __launch_bounds__(256) supcon_finish_kernel(const float4* __restrict__ stats,
                                                            const long long* __restrict__ lf,
                                                            const float* __restrict__ class_w, int N,
                                                            double* __restrict__ acc) {
  double t = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
    const float4 st = stats[i];
    const float lse = st.x + logf(st.y);
    float li = (st.z > 0.f ? lse : 0.f) - st.w / fmaxf(st.z, 1.0f);
    if (class_w) li *= __ldg(class_w + lf[i]);
    t += (double)li;
  }
  t = warp_sum_d(t);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, t);
}

// gradient w.r.t. the ROW embeddings: d f_i = sum_j g_ij a_j / T with
//   g_ij = coef_i * ( softmax_ij * (c_i > 0) - [label match] / max(c_i, 1) ),  coef_i = scale * w_i
// ROLE = 0: CTA owns rows i (outputs d f); ROLE = 1: CTA owns columns j (outputs d a), looping over the rows.
template <int D4, int ROLE>
__global__ void __launch_bounds__(256) supcon_bwd_kernel(const float* __restrict__ f, const float* __restrict__ a,
                                                         const long long* __restrict__ lf,
                                                         const long long* __restrict__ la, int N, int Na, int self_off,
                                                         float inv_t, const float4* __restrict__ stats,
                                                         const float* __restrict__ class_w,
                                                         const float* __restrict__ scale_dev, float* __restrict__ out) {
  // "own" = the side whose gradient this CTA produces; "oth" = the side it streams through shared memory
  __shared__ float4 s_o[SC_TJ][D4 + 1];
  __shared__ long long s_l[SC_TJ];
  __shared__ float4 s_st[SC_TJ];
  const int nown = ROLE == 0 ? N : Na, noth = ROLE == 0 ? Na : N;
  const float* own = ROLE == 0 ? f : a;
  const float* oth = ROLE == 0 ? a : f;
  const long long* lown = ROLE == 0 ? lf : la;
  const long long* loth = ROLE == 0 ? la : lf;
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  const int u = blockIdx.x * SC_TI + r;                      // own index
  const float sc = __ldg(scale_dev) * inv_t;
  float4 fu[D4], g[D4];
#pragma unroll
  for (int d = 0; d < D4; ++d) {
    fu[d] = (u < nown) ? __ldg(reinterpret_cast<const float4*>(own + (size_t)u * D4 * 4) + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    g[d] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long lu = (u < nown) ? lown[u] : -1;
  float4 stu = make_float4(0.f, 1.f, 0.f, 0.f);
  float wu = 1.0f;
  if (ROLE == 0 && u < N) { stu = stats[u]; wu = class_w ? __ldg(class_w + lu) : 1.0f; }
  for (int v0 = 0; v0 < noth; v0 += SC_TJ) {
    __syncthreads();
    for (int t = threadIdx.x; t < SC_TJ * D4; t += blockDim.x) {
      const int vv = t / D4, d = t - vv * D4;
      s_o[vv][d] = (v0 + vv < noth) ? __ldg(reinterpret_cast<const float4*>(oth + (size_t)(v0 + vv) * D4 * 4) + d)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x < SC_TJ) {
      const int v = v0 + threadIdx.x;
      s_l[threadIdx.x] = (v < noth) ? loth[v] : -2;
      if (ROLE == 1) {
        float4 st = make_float4(0.f, 1.f, 0.f, 0.f);
        if (v < N) { st = stats[v]; st.w = class_w ? __ldg(class_w + loth[v]) : 1.0f; }   // .w reused: row weight
        s_st[threadIdx.x] = st;
      }
    }
    __syncthreads();
    if (u < nown) {
      for (int vv = q; vv < SC_TJ; vv += 4) {
        const int v = v0 + vv;
        if (v >= noth) continue;
        const int i = ROLE == 0 ? u : v, j = ROLE == 0 ? v : u;
        if (j == i + self_off) continue;
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < D4; ++d) {
          const float4 ov = s_o[vv][d];
          dot = fmaf(fu[d].x, ov.x, dot); dot = fmaf(fu[d].y, ov.y, dot);
          dot = fmaf(fu[d].z, ov.z, dot); dot = fmaf(fu[d].w, ov.w, dot);
        }
        const float4 st = ROLE == 0 ? stu : s_st[vv];        // row statistics {m, s, c, (spos | w)}
        const float wi = ROLE == 0 ? wu : st.w;
        const float pij = st.z > 0.f ? expf(dot * inv_t - st.x) / st.y : 0.f;
        const float pos = (s_l[vv] == lu) ? 1.0f / fmaxf(st.z, 1.0f) : 0.f;
        const float gij = sc * wi * (pij - pos);
#pragma unroll
        for (int d = 0; d < D4; ++d) {
          const float4 ov = s_o[vv][d];
          g[d].x = fmaf(gij, ov.x, g[d].x); g[d].y = fmaf(gij, ov.y, g[d].y);
          g[d].z = fmaf(gij, ov.z, g[d].z); g[d].w = fmaf(gij, ov.w, g[d].w);
        }
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D4; ++d) {
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      g[d].x += __shfl_xor_sync(0xffffffffu, g[d].x, o); g[d].y += __shfl_xor_sync(0xffffffffu, g[d].y, o);
      g[d].z += __shfl_xor_sync(0xffffffffu, g[d].z, o); g[d].w += __shfl_xor_sync(0xffffffffu, g[d].w, o);
    }
    if (u < nown && q == 0) reinterpret_cast<float4*>(out + (size_t)u * D4 * 4)[d] = g[d];
  }
}

// y = x / max(||x||_2, eps)  (F.normalize) and its backward  dx = (dy - y <y, dy>) / max(||x||, eps)
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* __restrict__ x, int N, int D, float eps,
                                                          float* __restrict__ y, float* __restrict__ nrm) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < N; i += gridDim.x * wpb) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = x[(size_t)i * D + d]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float n = fmaxf(sqrtf(s), eps);
    for (int d = lane; d < D; d += 32) y[(size_t)i * D + d] = x[(size_t)i * D + d] / n;
    if (lane == 0) nrm[i] = n;
  }
}

__global__ void __launch_bounds__(256) l2norm_rows_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                              const float* __restrict__ nrm, int N, int D,
                                                              float* __restrict__ dx) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < N; i += gridDim.x * wpb) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(y[(size_t)i * D + d], dy[(size_t)i * D + d], s);
    s = warp_sum(s);
    const float inv = 1.0f / nrm[i];
    for (int d = lane; d < D; d += 32) dx[(size_t)i * D + d] = (dy[(size_t)i * D + d] - y[(size_t)i * D + d] * s) * inv;
  }
}

}  // namespace creste

using namespace creste;

extern "C" int creste_smooth_l1(const float* pred, const float* gt, const uint8_t* mask, long long n, float gt_scale,
                                float beta, double* acc2, void* stream) {
  CRESTE_CHECK_ARG(pred && gt && acc2 && n > 0 && beta > 0, "creste_smooth_l1: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(acc2, 0, 2 * sizeof(double), st));
  smooth_l1_kernel<<<l2_grid(n, 256, 148 * 8), 256, 0, st>>>(pred, gt, mask, n, gt_scale, beta, acc2);
  return launch_check("smooth_l1_kernel");
}

extern "C" int creste_smooth_l1_bwd(const float* pred, const float* gt, const uint8_t* mask, long long n,
                                    float gt_scale, float beta, const float* scale_dev, float* dpred, void* stream) {
  CRESTE_CHECK_ARG(pred && gt && scale_dev && dpred && n > 0 && beta > 0, "creste_smooth_l1_bwd: bad args");
  smooth_l1_bwd_kernel<<<l2_grid(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(pred, gt, mask, n, gt_scale, beta,
                                                                                  scale_dev, dpred);
  return launch_check("smooth_l1_bwd_kernel");
}

extern "C" int creste_ce_weighted(const float* logits_nchw, const int64_t* labels, const uint8_t* mask,
                                  const float* class_weights, int B, int C, long long HW, long long ignore_index,
                                  double* acc4, void* stream) {
  CRESTE_CHECK_ARG(logits_nchw && labels && acc4 && B > 0 && C > 0 && HW > 0, "creste_ce_weighted: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(acc4, 0, 4 * sizeof(double), st));
  ce_weighted_kernel<<<l2_grid((long long)B * HW, 256, 148 * 8), 256, 0, st>>>(
      logits_nchw, (const long long*)labels, mask, class_weights, B, C, HW, ignore_index, acc4);
  return launch_check("ce_weighted_kernel");
}

extern "C" int creste_ce_weighted_bwd(const float* logits_nchw, const int64_t* labels, const uint8_t* mask,
                                      const float* class_weights, int B, int C, long long HW, long long ignore_index,
                                      const float* scale_dev, float* dlogits, void* stream) {
  CRESTE_CHECK_ARG(logits_nchw && labels && scale_dev && dlogits && B > 0 && C > 0 && HW > 0,
                   "creste_ce_weighted_bwd: bad args");
  ce_weighted_bwd_kernel<<<l2_grid((long long)B * HW, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      logits_nchw, (const long long*)labels, mask, class_weights, B, C, HW, ignore_index, scale_dev, dlogits);
  return launch_check("ce_weighted_bwd_kernel");
}

extern "C" int creste_l2norm_rows(const float* x, int N, int D, float eps, float* y, float* norms, void* stream) {
  CRESTE_CHECK_ARG(x && y && norms && N > 0 && D > 0, "creste_l2norm_rows: bad args");
  l2norm_rows_kernel<<<l2_grid(N, 8, 148 * 8), 256, 0, (cudaStream_t)stream>>>(x, N, D, eps, y, norms);
  return launch_check("l2norm_rows_kernel");
}

extern "C" int creste_l2norm_rows_bwd(const float* y, const float* dy, const float* norms, int N, int D, float* dx,
                                      void* stream) {
  CRESTE_CHECK_ARG(y && dy && norms && dx && N > 0 && D > 0, "creste_l2norm_rows_bwd: bad args");
  l2norm_rows_bwd_kernel<<<l2_grid(N, 8, 148 * 8), 256, 0, (cudaStream_t)stream>>>(y, dy, norms, N, D, dx);
  return launch_check("l2norm_rows_bwd_kernel");
}

#define SUPCON_DISPATCH(D4v, CALL)            \
  switch (D4v) {                              \
    case 1: { constexpr int D4 = 1; CALL; break; }   \
    case 2: { constexpr int D4 = 2; CALL; break; }   \
    case 4: { constexpr int D4 = 4; CALL; break; }   \
    case 8: { constexpr int D4 = 8; CALL; break; }   \
    case 16: { constexpr int D4 = 16; CALL; break; } \
    case 32: { constexpr int D4 = 32; CALL; break; } \
    default: set_error("creste_supcon: D must be 4, 8, 16, 32, 64 or 128 (got %d)", 4 * (D4v)); return CRESTE_ERR_ARG; \
  }

extern "C" int creste_supcon_fwd(const float* f, const float* a, const int64_t* lf, const int64_t* la, int N, int Na,
                                 int D, int self_off, float temperature, const float* class_weights, float* stats4,
                                 double* loss_sum, void* stream) {
  CRESTE_CHECK_ARG(f && a && lf && la && stats4 && loss_sum && N > 0 && Na > 0 && D % 4 == 0 && temperature > 0,
                   "creste_supcon_fwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  CRESTE_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), st));
  const int grid = ceil_div(N, SC_TI);
  SUPCON_DISPATCH(D / 4, (supcon_rows_kernel<D4><<<grid, 256, 0, st>>>(f, a, (const long long*)lf, (const long long*)la, N,
                                                                      Na, self_off, 1.0f / temperature, (float4*)stats4)));
  int rc = launch_check("supcon_rows_kernel");
  if (rc) return rc;
  supcon_finish_kernel<<<l2_grid(N, 256, 148), 256, 0, st>>>((const float4*)stats4, (const long long*)lf, class_weights,
                                                            N, loss_sum);
  return launch_check("supcon_finish_kernel");
}

extern "C" int creste_supcon_bwd(const float* f, const float* a, const int64_t* lf, const int64_t* la, int N, int Na,
                                 int D, int self_off, float temperature, const float* class_weights,
                                 const float* stats4, const float* scale_dev, float* df, float* da, void* stream) {
  CRESTE_CHECK_ARG(f && a && lf && la && stats4 && scale_dev && df && da && N > 0 && Na > 0 && D % 4 == 0,
                   "creste_supcon_bwd: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const float it = 1.0f / temperature;
  SUPCON_DISPATCH(D / 4, (supcon_bwd_kernel<D4, 0><<<ceil_div(N, SC_TI), 256, 0, st>>>(
                             f, a, (const long long*)lf, (const long long*)la, N, Na, self_off, it,
                             (const float4*)stats4, class_weights, scale_dev, df)));
  int rc = launch_check("supcon_bwd_kernel<rows>");
  if (rc) return rc;
  SUPCON_DISPATCH(D / 4, (supcon_bwd_kernel<D4, 1><<<ceil_div(Na, SC_TI), 256, 0, st>>>(
                             f, a, (const long long*)lf, (const long long*)la, N, Na, self_off, it,
                             (const float4*)stats4, class_weights, scale_dev, da)));
  return launch_check("supcon_bwd_kernel<cols>");
}
