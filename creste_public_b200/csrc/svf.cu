// svf.cu -- expected state-visitation frequency (K15) + greedy rollout (K16).
//
// Replaces MaxEntIRL.expected_state_visitation_frequency (reference creste/models/lfd.py:156-277)
// and earliest_pose_in_fov (creste/utils/train_utils.py:765-803).  The reference runs T-1 = 49
// iterations of {mul, depthwise 3x3 conv with one-hot kernels, sum over actions} plus 49
// iterations of ~10 tiny indexing launches, and materialises mu [B,T,H*W].
//
// One CTA per sample.  mu_0 is a delta at S0 and mass moves at most one cell per step, so after
// t steps the support lies inside the (2t+1)^2 window around S0: the CTA keeps two ping-pong
// copies of that window (<= (2T-1)^2 cells, 39 KB each for T = 50) in shared memory, and the
// time-sum in registers.  The sharpened policy softmax((pi - max pi)/temperature) is computed
// once for the window into the caller's workspace (L2-resident), then every step is the gather
//      mu'(s) = sum_a  pol_a(s - d_a) * mu(s - d_a)          (a = 0..7 in order, zero off-grid)
// which is the reference's conv + sum(dim=1) in the same summation order.  The last warp does the
// 49-step arg-max rollout concurrently.  HBM traffic: pi read once (32 B/cell of the window),
// exp_svf written once -- far below the 40 B/cell/step the streamed reference moves.
#include "common.cuh"

namespace creste {

__constant__ int c_dyn[8][2] = {{-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};

struct SvfParams {
  const float* policy;      // [B,8,H,W]
  const float* expert_rc;   // [B,Te,2]
  const uint8_t* fov;       // [H,W]
  float* pol_ws;            // [B,8,Wh*Ww] sharpened policy of the window
  float* exp_svf;           // [B,H,W]   (zeroed by the host wrapper)
  long long* states;        // [B,T,2]
  float* states_grid;       // [B,H,W]   (zeroed by the host wrapper)
  int B, H, W, T, Te, ds, sharpen, zero_terminal;  // T: horizon, Te: expert poses
  int Wh, Ww;               // window size
  float temperature;
};

constexpr int SVF_THREADS = 512;
constexpr int SVF_MAX_CELLS_PER_THREAD = 24;  // (2*50-1)^2 / 480 = 20.4

__global__ void __launch_bounds__(SVF_THREADS) svf_kernel(SvfParams p) {
  extern __shared__ float sm[];  // mu ping-pong: 2 * (Wh+2)*(Ww+2) with a zero halo
  __shared__ int s_pose[4];      // s0r, s0c, s1r, s1c
  const int b = blockIdx.x;
  const int H = p.H, W = p.W, T = p.T, Te = p.Te;
  const int tid = threadIdx.x;
  const size_t HW = (size_t)H * W;
  const float* P = p.policy + (size_t)b * 8 * HW;

  if (tid == 0) {
    // S = (expert // ds).long() clamped (lfd.py:171-173); earliest pose inside the FOV mask,
    // else (H-1, W//2) (train_utils.py:765-803); S1 = last pose.
    int s0r = H - 1, s0c = W / 2, s1r = 0, s1c = 0;
    bool found = false;
    for (int t = 0; t < Te; ++t) {
      const float er = p.expert_rc[((size_t)b * Te + t) * 2 + 0];
      const float ec = p.expert_rc[((size_t)b * Te + t) * 2 + 1];
      long long rr = (long long)floorf(__fdiv_rn(er, (float)p.ds));
      long long cc = (long long)floorf(__fdiv_rn(ec, (float)p.ds));
      rr = rr < 0 ? 0 : (rr > H - 1 ? H - 1 : rr);
      cc = cc < 0 ? 0 : (cc > W - 1 ? W - 1 : cc);
      if (!found && p.fov[rr * W + cc]) { s0r = (int)rr; s0c = (int)cc; found = true; }
      if (t == Te - 1) { s1r = (int)rr; s1c = (int)cc; }
    }
    s_pose[0] = s0r; s_pose[1] = s0c; s_pose[2] = s1r; s_pose[3] = s1c;
  }
  __syncthreads();
  const int s0r = s_pose[0], s0c = s_pose[1], s1r = s_pose[2], s1c = s_pose[3];

  const int Wh = p.Wh, Ww = p.Ww;
  // window origin: S0 - (T-1), clamped so the window stays inside the grid
  int oy = s0r - (T - 1), ox = s0c - (T - 1);
  oy = max(0, min(oy, H - Wh));
  ox = max(0, min(ox, W - Ww));
  const int pitch = Ww + 2;
  const int plane = (Wh + 2) * pitch;
  float* mu0 = sm;
  float* mu1 = sm + plane;
  for (int i = tid; i < 2 * plane; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();

  const int ncell = Wh * Ww;
  const int nprop = blockDim.x - 32;  // the last warp runs the rollout, concurrently
  float* pol = p.pol_ws + (size_t)b * 8 * ncell;
  // barrier among the propagation warps only (named barrier 1)
  auto prop_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(nprop) : "memory"); };

  if (tid >= nprop) {
    // ---- greedy rollout on the ORIGINAL policy (lfd.py:230-248), one lane
    if (tid == nprop) {
      float* g = p.states_grid + (size_t)b * HW;
      long long* st = p.states + (size_t)b * T * 2;
      int cr = s0r, cc = s0c;
      st[0] = cr; st[1] = cc;
      g[(size_t)cr * W + cc] += 1.0f;
      for (int t = 1; t < T; ++t) {
        const size_t s = (size_t)cr * W + cc;
        float pv[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) pv[a] = __ldg(P + a * HW + s);
        int best = 0;
        float bv = pv[0];
#pragma unroll
        for (int a = 1; a < 8; ++a)
          if (pv[a] > bv) { bv = pv[a]; best = a; }
        cr = max(0, min(cr + c_dyn[best][0], H - 1));
        cc = max(0, min(cc + c_dyn[best][1], W - 1));
        st[t * 2] = cr; st[t * 2 + 1] = cc;
        g[(size_t)cr * W + cc] += 1.0f;
      }
    }
    return;
  }
  {
    // ---- sharpened policy of the window -> workspace (lfd.py:190-194)
    for (int i = tid; i < ncell; i += nprop) {
      const int wy = i / Ww, wx = i - wy * Ww;
      const size_t s = (size_t)(oy + wy) * W + (ox + wx);
      float pv[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) pv[a] = __ldg(P + a * HW + s);
      if (p.sharpen) {
        float m = pv[0];
#pragma unroll
        for (int a = 1; a < 8; ++a) m = fmaxf(m, pv[a]);
        float l[8], lm, e[8], sum = 0.0f;
#pragma unroll
        for (int a = 0; a < 8; ++a) l[a] = __fdiv_rn(__fsub_rn(pv[a], m), p.temperature);
        lm = l[0];
#pragma unroll
        for (int a = 1; a < 8; ++a) lm = fmaxf(lm, l[a]);
#pragma unroll
        for (int a = 0; a < 8; ++a) { e[a] = expf(__fsub_rn(l[a], lm)); sum = __fadd_rn(sum, e[a]); }
#pragma unroll
        for (int a = 0; a < 8; ++a) pv[a] = __fdiv_rn(e[a], sum);
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) pol[(size_t)a * ncell + i] = pv[a];
    }
  }
  __threadfence_block();
  prop_sync();  // window policy visible to the propagation warps

  float acc[SVF_MAX_CELLS_PER_THREAD];
#pragma unroll
  for (int j = 0; j < SVF_MAX_CELLS_PER_THREAD; ++j) acc[j] = 0.0f;
  if (tid == 0) mu0[(s0r - oy + 1) * pitch + (s0c - ox + 1)] = 1.0f;
  const bool s1_in = (s1r >= oy && s1r < oy + Wh && s1c >= ox && s1c < ox + Ww);
  const int s1_off = (s1r - oy + 1) * pitch + (s1c - ox + 1);
  prop_sync();

  float* cur = mu0;
  float* nxt = mu1;
  for (int t = 1; t < T; ++t) {
    if (p.zero_terminal) {  // mu[t-1][S1] = 0 before it is propagated and summed (lfd.py:202-203)
      if (tid == 0 && s1_in) cur[s1_off] = 0.0f;
      prop_sync();
    }
    {
      // after t - 1 steps the mass lies within t - 1 cells of S0, so only the cells within t of it can receive any:
      // everything outside this box is zero in both buffers and is skipped (adds of exact zeros; same results)
      const int by0 = s0r - oy - t, by1 = s0r - oy + t, bx0 = s0c - ox - t, bx1 = s0c - ox + t;
#pragma unroll
      for (int j = 0; j < SVF_MAX_CELLS_PER_THREAD; ++j) {
        const int i = tid + j * nprop;
        if (i < ncell) {
          const int wy = i / Ww, wx = i - wy * Ww;
          if (wy < by0 || wy > by1 || wx < bx0 || wx > bx1) continue;
          acc[j] = __fadd_rn(acc[j], cur[(wy + 1) * pitch + wx + 1]);
          float sum = 0.0f;
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            const int sy = wy - c_dyn[a][0], sx = wx - c_dyn[a][1];
            // source inside the window? (outside the window mu is zero by construction; the
            // grid border coincides with the window border whenever it matters)
            float c = 0.0f;
            if (sy >= 0 && sy < Wh && sx >= 0 && sx < Ww) {
              const float m = cur[(sy + 1) * pitch + sx + 1];
              if (m != 0.0f) c = __fmul_rn(pol[(size_t)a * ncell + sy * Ww + sx], m);
            }
            sum = __fadd_rn(sum, c);
          }
          nxt[(wy + 1) * pitch + wx + 1] = sum;
        }
      }
    }
    prop_sync();
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  {
    float* out = p.exp_svf + (size_t)b * HW;
#pragma unroll
    for (int j = 0; j < SVF_MAX_CELLS_PER_THREAD; ++j) {
      const int i = tid + j * nprop;
      if (i < ncell) {
        const int wy = i / Ww, wx = i - wy * Ww;
        out[(size_t)(oy + wy) * W + (ox + wx)] = __fadd_rn(acc[j], cur[(wy + 1) * pitch + wx + 1]);
      }
    }
  }
}

static void svf_window(int H, int W, int T, int* Wh, int* Ww) {
  const int side = 2 * (T - 1) + 1;
  *Wh = H < side ? H : side;
  *Ww = W < side ? W : side;
}

}  // namespace creste

using namespace creste;

extern "C" size_t creste_svf_workspace_bytes(int B, int H, int W, int T) {
  int Wh, Ww;
  svf_window(H, W, T, &Wh, &Ww);
  return (size_t)B * 8 * Wh * Ww * sizeof(float) + 256;
}

extern "C" int creste_svf(const float* policy, const float* expert_rc, const uint8_t* fov, int B,
                          int H, int W, int T, int T_expert, int ds, int sharpen, float temperature,
                          int zero_terminal_state, float* exp_svf, int64_t* states,
                          float* states_grid, void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(policy && expert_rc && fov && exp_svf && states && states_grid && ws,
                   "creste_svf: null pointer");
  CRESTE_CHECK_ARG(B > 0 && H > 0 && W > 0 && T > 0 && T_expert > 0 && ds > 0, "creste_svf: bad shape");
  if (ws_bytes < creste_svf_workspace_bytes(B, H, W, T)) {
    set_error("creste_svf: workspace too small");
    return CRESTE_ERR_WORKSPACE;
  }
  SvfParams p;
  svf_window(H, W, T, &p.Wh, &p.Ww);
  const int ncell = p.Wh * p.Ww;
  CRESTE_CHECK_ARG(ncell <= (SVF_THREADS - 32) * SVF_MAX_CELLS_PER_THREAD,
                   "creste_svf: window %dx%d too large for the resident kernel (T=%d)", p.Wh, p.Ww, T);
  const size_t smem = (size_t)2 * (p.Wh + 2) * (p.Ww + 2) * sizeof(float);
  CRESTE_CHECK_ARG(smem <= 200 * 1024, "creste_svf: window needs %zu B shared memory", smem);
  cudaStream_t st = (cudaStream_t)stream;
  p.policy = policy; p.expert_rc = expert_rc; p.fov = fov;
  p.pol_ws = (float*)ws;
  p.exp_svf = exp_svf; p.states = (long long*)states; p.states_grid = states_grid;
  p.B = B; p.H = H; p.W = W; p.T = T; p.Te = T_expert; p.ds = ds; p.sharpen = sharpen;
  p.zero_terminal = zero_terminal_state; p.temperature = temperature;
  const size_t n = (size_t)B * H * W * sizeof(float);
  CRESTE_CUDA(cudaMemsetAsync(exp_svf, 0, n, st));
  CRESTE_CUDA(cudaMemsetAsync(states_grid, 0, n, st));
  CRESTE_CUDA(cudaFuncSetAttribute(svf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  svf_kernel<<<B, SVF_THREADS, smem, st>>>(p);
  return launch_check("svf_kernel");
}
