// vi.cu -- value iteration (K14): one persistent cooperative launch per solve.
//
// Replaces VIN.value_iteration_manual (reference creste/models/blocks/vin.py:48-80, stencil
// weights :36-46).  The reference issues ~8 library launches + one .item() host sync per sweep
// (~690 sweeps); here the whole solve -- every sweep, the batch-global convergence test, and the
// final q / softmax-policy pass -- is a single launch with no host round trip.
//
// Layout: the B*H rows of the batch are one flat row space, cut into `G` contiguous strips of
// `R` rows, one CTA per strip (G <= #SMs so that all CTAs are co-resident; cooperative launch).
// Per sweep a CTA (1) builds X = r + gamma*v for its strip plus one halo row above/below in shared
// memory, (2) evaluates the eight 3-tap action stencils from the 3x3 window, takes the max,
// writes v' to the other ping-pong buffer and reduces max|v'-v| (warp shuffle -> shared ->
// one atomicMax per CTA), (3) passes one grid barrier (monotonic counter in global memory).
// After the barrier every CTA reads the same global delta and takes the same decision.
// v ping-pong buffers (2*B*H*W*4 bytes) stay L2-resident (<= 34 MB for B=64, 256x256).
//
// Arithmetic order is the reference's CPU order, bit for bit: X = fadd(r, fmul(v, gamma));
// q_a = fma(w3,x3, fma(w2,x2, fma(w1,x1, 0))) over the non-zero taps in (ky,kx) raster order
// (measured against torch CPU conv2d, DESIGN.md "Arithmetic order").  The sweep count K therefore
// equals the reference's.
//
// Roofline: algorithmic bytes = 12 B/cell/sweep (read r, read v, write v') + 76 B/cell for the
// final pass; HBM-bound if streamed, here served from L2/SMEM.
#include "common.cuh"

namespace creste {

struct ViTap { int dy, dx; float w; };
// per action: the three non-zero taps of vin.py:36-46 in (ky,kx) raster order
__constant__ ViTap c_vi_taps[8][3] = {
    {{-1, -1, 0.8f}, {-1, 0, 0.1f}, {0, -1, 0.1f}},
    {{-1, -1, 0.1f}, {-1, 0, 0.8f}, {-1, 1, 0.1f}},
    {{-1, 0, 0.1f}, {-1, 1, 0.8f}, {0, 1, 0.1f}},
    {{-1, -1, 0.1f}, {0, -1, 0.8f}, {1, -1, 0.1f}},
    {{-1, 1, 0.1f}, {0, 1, 0.8f}, {1, 1, 0.1f}},
    {{0, -1, 0.1f}, {1, -1, 0.8f}, {1, 0, 0.1f}},
    {{1, -1, 0.1f}, {1, 0, 0.8f}, {1, 1, 0.1f}},
    {{0, 1, 0.1f}, {1, 0, 0.1f}, {1, 1, 0.8f}},
};

struct ViParams {
  const float* r;
  float* va;
  float* vb;
  float* v_out;
  float* q_out;
  float* pi_out;
  unsigned* gdelta;    // [max_sweeps] float bits, zeroed by the host wrapper
  unsigned* counter;   // grid barrier, zeroed by the host wrapper
  int* sweeps_out;     // device int[2]
  int B, H, W, R, G, max_sweeps;
  float gamma, thr;
};

// 3x3 window of X around (lr, x) with the vertical validity of the *sample* applied
// (rows of neighbouring samples are adjacent in the flat row space).
__device__ __forceinline__ void load_window(const float* sm, int pitch, int lr, int x, bool up_ok,
                                            bool dn_ok, float win[3][3]) {
  const float* c = sm + (lr + 1) * pitch + (x + 1);
#pragma unroll
  for (int dx = -1; dx <= 1; ++dx) {
    win[0][dx + 1] = up_ok ? c[-pitch + dx] : 0.0f;
    win[1][dx + 1] = c[dx];
    win[2][dx + 1] = dn_ok ? c[pitch + dx] : 0.0f;
  }
}

__device__ __forceinline__ void eval_q(const float win[3][3], float q[8]) {
  // in-order FMA chains, written out so the compiler cannot re-associate
  // a0: (0,0)*.8, (0,1)*.1, (1,0)*.1
  q[0] = __fmaf_rn(0.1f, win[1][0], __fmaf_rn(0.1f, win[0][1], __fmul_rn(0.8f, win[0][0])));
  q[1] = __fmaf_rn(0.1f, win[0][2], __fmaf_rn(0.8f, win[0][1], __fmul_rn(0.1f, win[0][0])));
  q[2] = __fmaf_rn(0.1f, win[1][2], __fmaf_rn(0.8f, win[0][2], __fmul_rn(0.1f, win[0][1])));
  q[3] = __fmaf_rn(0.1f, win[2][0], __fmaf_rn(0.8f, win[1][0], __fmul_rn(0.1f, win[0][0])));
  q[4] = __fmaf_rn(0.1f, win[2][2], __fmaf_rn(0.8f, win[1][2], __fmul_rn(0.1f, win[0][2])));
  q[5] = __fmaf_rn(0.1f, win[2][1], __fmaf_rn(0.8f, win[2][0], __fmul_rn(0.1f, win[1][0])));
  q[6] = __fmaf_rn(0.1f, win[2][2], __fmaf_rn(0.8f, win[2][1], __fmul_rn(0.1f, win[2][0])));
  q[7] = __fmaf_rn(0.8f, win[2][2], __fmaf_rn(0.1f, win[2][1], __fmul_rn(0.1f, win[1][2])));
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*((volatile unsigned*)counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024, 1) vi_persistent_kernel(ViParams p) {
  extern __shared__ float sm[];  // X tile [(R+2)][W+2]
  __shared__ float s_red[32];
  __shared__ float s_delta;
  const int W = p.W, H = p.H, BH = p.B * p.H;
  const int pitch = W + 2;
  const int row0 = blockIdx.x * p.R;
  const int row1 = min(row0 + p.R, BH);
  const int n = row1 - row0;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float gamma = p.gamma;

  // halo columns are zero forever
  for (int i = tid; i < n + 2; i += nt) {
    sm[i * pitch] = 0.0f;
    sm[i * pitch + W + 1] = 0.0f;
  }

  auto build_x = [&](const float* vin, bool v_is_zero) {
    for (int idx = tid; idx < (n + 2) * W; idx += nt) {
      const int lr = idx / W - 1, x = idx - (lr + 1) * W;
      const int g = row0 + lr;
      float val = 0.0f;
      if (g >= 0 && g < BH) {
        const size_t o = (size_t)g * W + x;
        const float rr = __ldg(p.r + o);
        const float vv = v_is_zero ? 0.0f : __ldcg(vin + o);
        val = __fadd_rn(rr, __fmul_rn(vv, gamma));
      }
      sm[(lr + 1) * pitch + x + 1] = val;
    }
  };

  int K = 0;
  int hit_max = 1;
  const float* vfin = p.va;
  for (int k = 0; k < p.max_sweeps; ++k) {
    const float* vin = (k & 1) ? p.vb : p.va;
    float* vout = (k & 1) ? p.va : p.vb;
    build_x(vin, k == 0);
    __syncthreads();
    float dmax = 0.0f;
    for (int idx = tid; idx < n * W; idx += nt) {
      const int lr = idx / W, x = idx - lr * W;
      const int g = row0 + lr;
      const int y = g % H;
      float win[3][3], q[8];
      load_window(sm, pitch, lr, x, y > 0, y < H - 1, win);
      eval_q(win, q);
      float m = q[0];
#pragma unroll
      for (int a = 1; a < 8; ++a) m = fmaxf(m, q[a]);
      const size_t o = (size_t)g * W + x;
      const float vold = (k == 0) ? 0.0f : __ldcg(vin + o);
      dmax = fmaxf(dmax, fabsf(__fsub_rn(m, vold)));
      __stcg(vout + o, m);
    }
    dmax = warp_max(dmax);
    if ((tid & 31) == 0) s_red[tid >> 5] = dmax;
    __syncthreads();
    if (tid < 32) {
      float d = (tid < (nt >> 5)) ? s_red[tid] : 0.0f;
      d = warp_max(d);
      if (tid == 0) atomicMax(p.gdelta + k, __float_as_uint(d));
    }
    grid_barrier(p.counter, (unsigned)(k + 1) * (unsigned)p.G);
    if (tid == 0) s_delta = __uint_as_float(__ldcg(p.gdelta + k));
    __syncthreads();
    K = k + 1;
    vfin = vout;
    if (!(s_delta > p.thr)) {
      hit_max = 0;
      break;
    }
  }

  // final pass: q = conv(r + gamma*v), pi = softmax_a(q)   (vin.py:76-80)
  build_x(vfin, K == 0);
  __syncthreads();
  const size_t HW = (size_t)H * W;
  for (int idx = tid; idx < n * W; idx += nt) {
    const int lr = idx / W, x = idx - lr * W;
    const int g = row0 + lr;
    const int b = g / H, y = g - b * H;
    const size_t o = (size_t)g * W + x;
    if (p.v_out) p.v_out[o] = (K == 0) ? 0.0f : __ldcg(vfin + o);
    if (p.q_out || p.pi_out) {
      float win[3][3], q[8], e[8];
      load_window(sm, pitch, lr, x, y > 0, y < H - 1, win);
      eval_q(win, q);
      float m = q[0];
#pragma unroll
      for (int a = 1; a < 8; ++a) m = fmaxf(m, q[a]);
      float s = 0.0f;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        e[a] = expf(__fsub_rn(q[a], m));
        s = __fadd_rn(s, e[a]);
      }
      const size_t qo = (size_t)b * 8 * HW + (size_t)y * W + x;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        if (p.q_out) p.q_out[qo + a * HW] = q[a];
        if (p.pi_out) p.pi_out[qo + a * HW] = __fdiv_rn(e[a], s);
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0 && p.sweeps_out) {
    p.sweeps_out[0] = K;
    p.sweeps_out[1] = hit_max;
  }
}

// rows per CTA: as many CTAs as there are SMs, but never less than 2 rows per CTA, and a
// single CTA (no grid barrier traffic) when the whole problem is tiny.
static void vi_partition(int B, int H, int W, int* R, int* G) {
  const int BH = B * H;
  const int sms = num_sms() > 0 ? num_sms() : 148;
  if ((long long)BH * W <= 8192) {
    *R = BH;
    *G = 1;
    return;
  }
  int r = ceil_div(BH, sms);
  if (r < 2) r = 2;
  *R = r;
  *G = ceil_div(BH, r);
}

}  // namespace creste

using namespace creste;

extern "C" size_t creste_vi_workspace_bytes(int B, int H, int W, int max_sweeps) {
  const size_t n = (size_t)B * H * W;
  return 2 * align_up(n * sizeof(float), 256) + align_up((size_t)(max_sweeps + 1) * 4, 256) + 256;
}

extern "C" int creste_vi_solve(const float* r, float* v_out, float* q_out, float* pi_out, int B,
                               int H, int W, float gamma, float thr, int max_sweeps,
                               int* sweeps_out, void* ws, size_t ws_bytes, void* stream) {
  CRESTE_CHECK_ARG(r && ws, "creste_vi_solve: null r/ws");
  CRESTE_CHECK_ARG(B > 0 && H > 0 && W > 0 && max_sweeps > 0, "creste_vi_solve: bad shape");
  if (ws_bytes < creste_vi_workspace_bytes(B, H, W, max_sweeps)) {
    set_error("creste_vi_solve: workspace %zu < %zu", ws_bytes,
              creste_vi_workspace_bytes(B, H, W, max_sweeps));
    return CRESTE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * H * W;
  char* w = (char*)ws;
  ViParams p;
  p.r = r;
  p.va = (float*)w;
  w += align_up(n * sizeof(float), 256);
  p.vb = (float*)w;
  w += align_up(n * sizeof(float), 256);
  p.gdelta = (unsigned*)w;
  const size_t gbytes = align_up((size_t)(max_sweeps + 1) * 4, 256);
  w += gbytes;
  p.counter = (unsigned*)w;
  p.v_out = v_out;
  p.q_out = q_out;
  p.pi_out = pi_out;
  p.sweeps_out = sweeps_out;
  p.B = B; p.H = H; p.W = W;
  p.max_sweeps = max_sweeps;
  p.gamma = gamma; p.thr = thr;
  vi_partition(B, H, W, &p.R, &p.G);
  const size_t smem = (size_t)(p.R + 2) * (W + 2) * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("creste_vi_solve: strip of %d rows x %d cols needs %zu B shared memory", p.R, W, smem);
    return CRESTE_ERR_ARG;
  }
  CRESTE_CUDA(cudaMemsetAsync(p.gdelta, 0, gbytes + 256, st));
  CRESTE_CUDA(cudaFuncSetAttribute(vi_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  const long long cells = (long long)p.R * W;
  const int threads = cells <= 1024 ? 256 : (cells <= 8192 ? 512 : 1024);
  void* args[] = {&p};
  if (p.G > 1) {
    CRESTE_CUDA(cudaLaunchCooperativeKernel((void*)vi_persistent_kernel, dim3(p.G), dim3(threads),
                                            args, smem, st));
  } else {
    vi_persistent_kernel<<<1, threads, smem, st>>>(p);
  }
  return launch_check("vi_persistent_kernel");
}
