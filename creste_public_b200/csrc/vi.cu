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
#include <cooperative_groups.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace creste {

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


// ---------------------------------------------------------------------------------------------
// Cluster-resident variant: one thread-block cluster per sample, v and r in REGISTERS, the
// X = r + gamma*v tile of each CTA in shared memory (double buffered by sweep parity), halo rows
// pushed into the neighbour CTA's tile through distributed shared memory, ONE cluster barrier per
// sweep, and NO grid barrier: the batch-global max|dv| is posted fire-and-forget to global memory
// and consumed two sweeps later (three generations of v are kept in registers so that the solve
// can roll back to exactly the sweep the reference stops at -- K and v stay bit-identical).
namespace cg = cooperative_groups;

struct ViClusterParams {
  const float* r; float* v_out; float* q_out; float* pi_out;
  unsigned* gdelta;    // [max_sweeps] float bits (zeroed)
  unsigned* garrive;   // [max_sweeps] CTA arrival counts (zeroed)
  int* sweeps_out;
  int B, H, W, R, c, max_sweeps, G;
  float gamma, thr;
};

constexpr int VI_RING = 8;   // smem slots for per-sweep block maxima (>= LAG + 2)

__device__ __forceinline__ uint32_t vi_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void vi_mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(vi_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void vi_mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(vi_smem_u32(bar)), "r"(rank)
      : "memory");
}
__device__ __forceinline__ void vi_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "VI_WAIT:\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra VI_DONE;\n\t"
      "bra VI_WAIT;\n\t"
      "VI_DONE:\n\t"
      "}" ::"r"(vi_smem_u32(bar)), "r"(parity)
      : "memory");
}

// LAG: how many sweeps the global stop decision may trail the computation.  LAG + 2 generations
// of v live in registers so the solve rolls back to exactly the reference's stopping sweep.
// Synchronisation per sweep: every compute thread arrives on its CTA's mbarrier after writing its
// X values; a thread that pushes a halo value into a neighbour CTA also arrives (release.cluster)
// on the neighbour's mbarrier.  A CTA starts the stencil when its own writes and both neighbours'
// halo rows have landed -- no cluster-wide or grid-wide barrier in the loop.
template <int CELLS, int LAG>
__global__ void __launch_bounds__(544, 1) vi_cluster_kernel(ViClusterParams p) {
  extern __shared__ float sm[];   // two X tiles [(R+2)][W+2]
  __shared__ uint64_t s_mbar[2];
  __shared__ unsigned s_blockmax[VI_RING];
  __shared__ unsigned s_count[VI_RING];
  __shared__ volatile int s_dec[VI_RING];  // per sweep slot: (sweep << 1) | stop, published by the comm warp
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();           // strip index inside the sample
  const int b = blockIdx.x / p.c;
  const int W = p.W, H = p.H;
  const int pitch = W + 2;
  const int tile = (p.R + 2) * pitch;
  const int row0 = cr * p.R;
  const int n = min(p.R, H - row0);                    // rows of this strip (>= 1 by construction)
  const int tid = threadIdx.x;
  const int nt = blockDim.x - 32;                      // compute threads; the last warp is the comm warp
  const int nwarps = nt >> 5;
  const int ncell = n * W;
  const float gamma = p.gamma;
  const size_t base = ((size_t)b * H + row0) * W;
  const bool is_comm = tid >= nt;
  const bool has_up = cr > 0;
  const bool has_dn = (cr < p.c - 1) && (row0 + n < H);

  for (int i = tid; i < 2 * tile; i += blockDim.x) sm[i] = 0.0f;   // halos (incl. sample borders) stay 0
  if (tid < VI_RING) { s_blockmax[tid] = 0u; s_count[tid] = 0u; }
  if (tid < VI_RING) s_dec[tid] = -2;
  if (tid == 0) {
    const unsigned cnt = (unsigned)nt + (has_up ? (unsigned)W : 0u) + (has_dn ? (unsigned)W : 0u);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vi_smem_u32(&s_mbar[0])), "r"(cnt) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vi_smem_u32(&s_mbar[1])), "r"(cnt) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float* up_tile = has_up ? cluster.map_shared_rank(sm, cr - 1) : nullptr;
  float* dn_tile = has_dn ? cluster.map_shared_rank(sm, cr + 1) : nullptr;
  const int n_up = p.R;                                  // every strip above a strip is full
  cluster.sync();                                        // zero-fill + barrier init visible cluster-wide

  if (is_comm) {
    // ===== communication warp: posts this CTA's per-sweep max|dv| to global memory and turns the
    // batch-global maxima into stop decisions, off the compute warps' critical path.  Lane l serves
    // sweeps k = l, l + VI_RING, ...: up to VI_RING sweeps are in flight, so the ~1-2 us global round
    // trip bounds the decision LATENCY (absorbed by LAG), not the sweep rate. =====
    // Written as a NON-BLOCKING state machine stepped in lock-step by the whole warp: a lane never
    // spins inside divergent code (a spinning lane would hold the others at the reconvergence point
    // and deadlock the decision pipeline).
    const int lane = tid - nt;
    int k = lane;                       // sweep this lane is serving
    int state = (lane < VI_RING && k < p.max_sweeps) ? 0 : 2;   // 0: wait local, 1: wait global, 2: done
    while (__any_sync(0xffffffffu, state != 2)) {
      if (state != 2) {
        // an earlier sweep was decided as the last one -> nothing more will be posted
        for (int j = 0; j < VI_RING; ++j) {
          const int dj = s_dec[j];
          if (dj >= 0 && (dj & 1) && (dj >> 1) < k) state = 2;
        }
      }
      if (state == 0) {
        const int slot = k % VI_RING;
        if (*((volatile unsigned*)&s_count[slot]) >= (unsigned)nwarps) {
          __threadfence_block();
          const unsigned d = *((volatile unsigned*)&s_blockmax[slot]);
          s_blockmax[slot] = 0u;
          s_count[slot] = 0u;
          __threadfence_block();
          atomicMax(p.gdelta + k, d);
          __threadfence();
          atomicAdd(p.garrive + k, 1u);
          state = 1;
        }
      } else if (state == 1) {
        if (*((volatile unsigned*)(p.garrive + k)) >= (unsigned)p.G) {
          __threadfence();
          const float dk = __uint_as_float(*((volatile unsigned*)(p.gdelta + k)));
          const int stop = !(dk > p.thr);
          __threadfence_block();
          s_dec[k % VI_RING] = (k << 1) | stop;
          k += VI_RING;
          state = (stop || k >= p.max_sweeps) ? 2 : 0;
        }
      }
    }
    __syncwarp();
  } else {
    int K = p.max_sweeps, hit_max = 1;
    float vf[CELLS];
    int off[CELLS];
    float rr[CELLS];
    float vh[LAG + 2][CELLS];   // vh[0] = scratch / newest, vh[1] = v_k at the top of sweep k, ...
#pragma unroll
    for (int j = 0; j < CELLS; ++j) {
      const int idx = tid + j * nt;
      const int lr = idx / W, x = idx - lr * W;
      off[j] = idx < ncell ? (lr + 1) * pitch + x + 1 : -1;
      rr[j] = idx < ncell ? __ldg(p.r + base + idx) : 0.0f;
#pragma unroll
      for (int g = 0; g < LAG + 2; ++g) vh[g][j] = 0.0f;
    }
    // write X = r + gamma*v for the own cells (+ halo pushes) into tile `idx & 1`, then arrive
    auto push_x = [&](const float (&v)[CELLS], int idx) {
      const int buf = idx & 1;
      float* mine = sm + buf * tile;
      uint64_t* bar = &s_mbar[buf];
#pragma unroll
      for (int j = 0; j < CELLS; ++j) {
        if (off[j] >= 0) {
          const float X = __fadd_rn(rr[j], __fmul_rn(v[j], gamma));
          mine[off[j]] = X;
          const int lr = off[j] / pitch - 1;
          if (lr == 0 && has_up) {
            up_tile[buf * tile + off[j] + n_up * pitch] = X;
            vi_mbar_arrive_remote(bar, (uint32_t)(cr - 1));
          }
          if (lr == n - 1 && has_dn) {
            dn_tile[buf * tile + off[j] - n * pitch] = X;
            vi_mbar_arrive_remote(bar, (uint32_t)(cr + 1));
          }
        }
      }
      vi_mbar_arrive_local(bar);
    };
    auto wait_x = [&](int idx) { vi_mbar_wait(&s_mbar[idx & 1], (uint32_t)((idx >> 1) & 1)); };

    int k = 0;
    int stop_at = -1;
    for (; k < p.max_sweeps; ++k) {
      push_x(vh[1], k);
      wait_x(k);
      const float* X = sm + (k & 1) * tile;
      float dmax = 0.0f;
#pragma unroll
      for (int j = 0; j < CELLS; ++j) {
        if (off[j] >= 0) {
          const float* c = X + off[j];
          float win[3][3], q[8];
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            win[0][dx + 1] = c[-pitch + dx];
            win[1][dx + 1] = c[dx];
            win[2][dx + 1] = c[pitch + dx];
          }
          eval_q(win, q);
          float m = q[0];
#pragma unroll
          for (int a = 1; a < 8; ++a) m = fmaxf(m, q[a]);
          vh[0][j] = m;
          dmax = fmaxf(dmax, fabsf(__fsub_rn(m, vh[1][j])));
        }
      }
      dmax = warp_max(dmax);
      if ((tid & 31) == 0) {
        const int slot = k % VI_RING;
        atomicMax(&s_blockmax[slot], __float_as_uint(dmax));
        __threadfence_block();
        atomicAdd(&s_count[slot], 1u);
      }
      // the decision of sweep k - LAG must be known before the generation it would select is dropped
      if (k >= LAG) {
        const int j = k - LAG;
        int dj;
        while (((dj = s_dec[j % VI_RING]) >> 1) != j) {
        }
        if (dj & 1) { stop_at = j; break; }
      }
#pragma unroll
      for (int g = LAG + 1; g >= 1; --g)
#pragma unroll
        for (int j = 0; j < CELLS; ++j) vh[g][j] = vh[g - 1][j];
    }
    //  break at sweep k with stop_at = j = k - LAG (arrays not rotated): vh[0] = v_{k+1},
    //  vh[g] = v_{k+1-g}; the answer v_{j+1} = vh[LAG].  Loop exhausted (rotated): vh[1] = v_ms,
    //  vh[g] = v_{ms+1-g}; the pending decisions j = ms-LAG .. ms-1 are resolved in order.
    int gen = 1;
    int next_idx = k;            // next unused tile / mbarrier phase index
    if (stop_at >= 0) {
      gen = LAG; K = stop_at + 1; hit_max = 0; next_idx = k + 1;
    } else {
      for (int j = max(0, p.max_sweeps - LAG); j < p.max_sweeps; ++j) {
        int dj;
        while (((dj = s_dec[j % VI_RING]) >> 1) != j) {
        }
        if (dj & 1) { K = j + 1; hit_max = 0; gen = p.max_sweeps - j; break; }
      }
    }
#pragma unroll
    for (int j = 0; j < CELLS; ++j) {
      float val = vh[1][j];
#pragma unroll
      for (int g = 0; g < LAG + 2; ++g)
        if (g == gen) val = vh[g][j];
      vf[j] = val;
    }
    // final pass on the selected generation: rebuild X, exchange halos, q / softmax (vin.py:76-80)
    push_x(vf, next_idx);
    wait_x(next_idx);
    const float* X = sm + (next_idx & 1) * tile;
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int j = 0; j < CELLS; ++j) {
      if (off[j] >= 0) {
        const int idx = tid + j * nt;
        const int lr = idx / W, x = idx - lr * W;
        const int y = row0 + lr;
        if (p.v_out) p.v_out[base + idx] = vf[j];
        if (p.q_out || p.pi_out) {
          const float* c = X + off[j];
          float win[3][3], q[8], e[8];
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            win[0][dx + 1] = c[-pitch + dx];
            win[1][dx + 1] = c[dx];
            win[2][dx + 1] = c[pitch + dx];
          }
          eval_q(win, q);
          float m = q[0];
#pragma unroll
          for (int a = 1; a < 8; ++a) m = fmaxf(m, q[a]);
          float s = 0.0f;
#pragma unroll
          for (int a = 0; a < 8; ++a) { e[a] = expf(__fsub_rn(q[a], m)); s = __fadd_rn(s, e[a]); }
          const size_t qo = (size_t)b * 8 * HW + (size_t)y * W + x;
#pragma unroll
          for (int a = 0; a < 8; ++a) {
            if (p.q_out) p.q_out[qo + a * HW] = q[a];
            if (p.pi_out) p.pi_out[qo + a * HW] = __fdiv_rn(e[a], s);
          }
        }
      }
    }
    if (blockIdx.x == 0 && tid == 0 && p.sweeps_out) {
      p.sweeps_out[0] = K;
      p.sweeps_out[1] = hit_max;
    }
  }
  cluster.sync();   // no CTA may exit while a neighbour can still write into its shared memory
}

// Try to launch the cluster-resident kernel; returns 1 if launched, 0 if not applicable, <0 / >0 on error.
template <int CELLS, int LAG>
static int vi_try_cluster(ViClusterParams& p, int c, int threads, size_t smem, cudaStream_t st) {
  auto kern = vi_cluster_kernel<CELLS, LAG>;
  threads += 32;   // + the communication warp
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (c > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.B * c);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (getenv("CRESTE_VI_DEBUG"))
    fprintf(stderr, "[creste_vi] B=%d H=%d W=%d c=%d R=%d CELLS=%d threads=%d smem=%zu max_clusters=%d\n",
            p.B, p.H, p.W, c, p.R, CELLS, threads, smem, max_clusters);
  if (max_clusters < p.B) return 0;      // every cluster must be co-resident (global delta exchange)
  p.c = c; p.G = p.B * c;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  count_launch();
  return 1;
}

// ---------------------------------------------------------------------------------------------
// Cluster-resident kernel, second generation ("strip" kernel): same data placement as above
// (one cluster per sample, r / v in registers, X tiles in shared memory, DSMEM halo exchange, no
// grid barrier) with the three costs measured on the first generation removed:
//   * register-blocked cells: a thread owns 4 adjacent columns x RT adjacent rows and reads each X
//     row with one LDS.128 + 2 scalar loads (1.5-2 smem loads per cell instead of 9);
//   * halo rows travel as st.async.v4 stores that complete transaction bytes on the NEIGHBOUR's
//     mbarrier (one 16-byte store per 4 cells, no per-cell remote arrive);
//   * the batch-global stop decision trails the computation by exactly LAGCHK sweeps (the ~2 us L2
//     round trip of the max-reduction is off the critical path) and bit-exactness of K and v is
//     kept by CHECKPOINT + REPLAY instead of rotating LAG+2 generations of v through registers:
//     v is snapshotted every CKPT sweeps (two snapshots live); once sweep K is known to be the
//     reference's last, every CTA restores the snapshot at floor(K/CKPT)*CKPT and replays
//     K - that many sweeps (same arithmetic, same order => same bits).  All CTAs observe the
//     decision at the same sweep index, so the whole grid stays in lock step without a barrier.
// Cluster size is any 1..16 (largest that is co-resident for the whole batch).
constexpr int VI2_RING = 32;     // per-sweep block-max slots in flight (one comm-warp lane each)
// The decision for sweep s is consumed at sweep s + lag, and v is snapshotted every `lag` sweeps
// (ViStripParams::lag, <= RING - 1): 8 for large strips (a sweep takes > 1 us, the decision ~4 us),
// 24 for small ones, where the sweep rate would otherwise be decision-latency / lag.
constexpr unsigned VI2_SPIN_LIMIT = 1u << 27;   // bail out instead of hanging the GPU

struct ViStripParams {
  const float* r; float* v_out; float* q_out; float* pi_out;
  unsigned long long* gword;   // [max_sweeps + 2] : low 32 = max|dv| bits, high 32 = arrival count
  int* sweeps_out;
  int B, H, W, R, c, max_sweeps, G, nt, lag;   // nt = compute threads (multiple of 32)
  int rgn, nfull;              // row groups per strip; the LAST nfull groups own RT rows, the others RT - 1
  float gamma, thr;
};

__device__ __forceinline__ uint32_t vi_mapa(uint32_t addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ void vi_st_async_v4(uint32_t raddr, float4 v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(raddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w)), "r"(rbar)
               : "memory");
}
// CTA-scope semantics on purpose: the tile is read only by this CTA's threads, and the neighbours'
// halo rows arrive through the async proxy (st.async ... complete_tx), whose data is visible to
// whoever observes the phase completion -- the TMA-multicast pattern.  Cluster-scope release /
// acquire compiled to MEMBAR.ALL.GPU + CCTL.IVALL per thread per sweep (measured: 2x slower).
__device__ __forceinline__ void vi_mbar_arrive_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vi_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void vi_mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(vi_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool vi_mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(vi_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// the same on 32-bit shared-window addresses held in registers (the hot loop sets every address up once)
__device__ __forceinline__ void vi_mbar_arrive_expect_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vi_mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool vi_mbar_test_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// shared-memory accesses at [register + compile-time byte offset]: with a compile-time tile pitch every row of the
// window is one base register plus an immediate, no address arithmetic in the loop
template <int OFF>
__device__ __forceinline__ void vi_sts_v4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0+%5], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "n"(OFF) : "memory");
}
__device__ __forceinline__ void vi_sts_u32(uint32_t a, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ float4 vi_lds_v4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ float vi_lds(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
  return v;
}
// compile-time loop: f(std::integral_constant<int, I>) for I = 0 .. N-1
template <int I, int N, class F>
__device__ __forceinline__ void vi_static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    vi_static_for<I + 1, N>(f);
  }
}
// three-input maximum (FMNMX3): exact, so any association of a max tree gives the same bits
__device__ __forceinline__ float vi_max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// PW: compile-time tile pitch in floats (W + 8) for the widths that matter (64, 128, 256), 0 = run-time pitch
template <int RT, int PW>
__global__ void __launch_bounds__(RT <= 2 ? 896 : (RT == 3 ? 608 : 544), 1) vi_strip_kernel(ViStripParams p) {
  // two X tiles [(R+2)][W+8] (interior col x at 4+x), then two v snapshots [R][W]
  extern __shared__ __align__(16) float sm[];
  __shared__ uint64_t s_mbar[2];
  __shared__ uint64_t s_postbar[VI2_RING];   // per-sweep "all warps posted their max|dv|" barriers
  __shared__ unsigned s_wmax[VI2_RING][32];  // per-sweep, per-warp max|dv| (float bits)
  __shared__ volatile int s_done;            // every sweep <= s_done is decided (or lies beyond a known stop)
  __shared__ volatile int s_stopK;           // smallest sweep decided as the last one
  __shared__ volatile int s_fail;
  cg::cluster_group cluster = cg::this_cluster();
  const int cr = (int)cluster.block_rank();
  const int b = blockIdx.x / p.c;
  const int W = p.W, H = p.H;
  const int P = PW ? PW : W + 8;
  const int tile = (p.R + 2) * P;
  const int row0 = cr * p.R;
  const int n = min(p.R, H - row0);
  const int tid = threadIdx.x;
  const int nt = p.nt;
  const int nwarps = nt >> 5;
  const bool is_comm = tid >= nt;
  const bool has_up = cr > 0;
  const bool has_dn = (cr < p.c - 1) && (row0 + n < H);
  const float gamma = p.gamma;

  // halos / sample borders stay 0; snapshot slot 0 starts as v_0 = 0
  for (int i = tid; i < 2 * tile + 2 * p.R * W; i += blockDim.x) sm[i] = 0.0f;
  if (tid == 0) {
    s_done = 0;
    s_stopK = 0x7fffffff;
    s_fail = 0;
    for (int i = 0; i < VI2_RING; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vi_smem_u32(&s_postbar[i])), "r"(nwarps) : "memory");
    const int arrivals = nwarps;          // one arrival per compute warp per sweep
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vi_smem_u32(&s_mbar[0])), "r"(arrivals) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vi_smem_u32(&s_mbar[1])), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();

  if (is_comm) {
    // ===== communication warp: lane l serves sweeps l+1, l+1+RING, ... (non-blocking state machine)
    const int lane = tid - nt;
    int s = lane + 1;
    int state = (lane < VI2_RING && s <= p.max_sweeps) ? 0 : 2;   // 0 wait local, 1 wait global, 2 done
    unsigned spins = 0;
    int done_pub = 0;
    while (__any_sync(0xffffffffu, state != 2)) {
      if (state != 2 && (s > s_stopK || s_fail)) state = 2;
      if (state == 0) {
        const int slot = (s - 1) % VI2_RING;
        if (vi_mbar_test(&s_postbar[slot], (uint32_t)(((s - 1) / VI2_RING) & 1))) {
          unsigned d = 0u;
          for (int wi = 0; wi < nwarps; ++wi) d = max(d, *((volatile unsigned*)&s_wmax[slot][wi]));
          unsigned* w = reinterpret_cast<unsigned*>(p.gword + s);
          asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(w), "r"(d) : "memory");
          asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(w + 1), "r"(1u) : "memory");
          state = 1;
        }
      } else if (state == 1) {
        unsigned long long word;
        asm volatile("ld.acquire.gpu.global.b64 %0, [%1];" : "=l"(word) : "l"(p.gword + s) : "memory");
        if ((unsigned)(word >> 32) >= (unsigned)p.G) {
          const float dk = __uint_as_float((unsigned)(word & 0xffffffffu));
          const int stop = !(dk > p.thr);
          if (stop) atomicMin((int*)&s_stopK, s);
          s += VI2_RING;
          state = (stop || s > p.max_sweeps) ? 2 : 0;
        }
      }
      // watermark: every sweep below the smallest undecided one is decided.  A lane that retired (past max_sweeps,
      // or beyond a stop that is already in s_stopK) no longer holds the watermark back.
      __threadfence_block();
      const int low = __reduce_min_sync(0xffffffffu, state == 2 ? 0x7fffffff : s);
      if (lane == 0 && low - 1 > done_pub) { done_pub = low - 1; s_done = done_pub; }
      if (++spins > VI2_SPIN_LIMIT) { s_fail = 1; break; }
      __nanosleep(40);      // leave the issue slots of this scheduler to the compute warps
    }
    __syncwarp();
  } else {
    // ===== compute threads: 4 columns x RT rows each.  The hot loop is branch-free apart from warp-uniform
    // branches: every thread loads and evaluates its RT rows unconditionally (rows past the strip read in-bounds
    // junk and keep junk in their v registers, which nothing ever stores) and only the tile stores / the running
    // maximum are predicated by the per-row `on` mask.  Round-2 rewrite: the integer / control work of the loop
    // (it ran on the 16-lane ALU pipe and cost more cycles than the FP work, profiles/r2_vi_experiments.md) is
    // hoisted: shared-memory addresses are 32-bit registers set up once, the sweeps come in blocks of `lag` with
    // a compile-time tile parity, and the stop decision is consumed once per block instead of once per sweep.
    const int cgn = W >> 2;
    const int rgn = p.rgn;
    const bool thread_on = tid < cgn * rgn;
    const int cgi = thread_on ? tid % cgn : 0, rg = thread_on ? tid / cgn : 0;
    // Row groups of RT - 1 rows first, then p.nfull groups of RT rows (the host only mixes the two when a warp never
    // spans two groups, so the row count is warp-uniform): with W = 256, R = 26 this is 6 x 3 + 2 x 4 rows on 16
    // compute warps = 13 rows per scheduler, where 9 uniform groups of 3 on 18 warps put 15 (one of them junk) on two of
    // the four schedulers.
    const int nshort = rgn - p.nfull;
    const int my_rows = rg >= nshort ? RT : RT - 1;
    const int lr0 = rg * (RT - 1) + max(0, rg - nshort);
    const int x0 = cgi << 2;
    bool failed = false;
    int K = p.max_sweeps, hit_max = 1;
    auto body = [&](auto nr_c) {
    constexpr int NR = decltype(nr_c)::value;
    bool on[NR];
    float4 rr[NR], v[NR];
    float* ckpt = sm + 2 * tile;                 // snapshot slot q at ckpt + q * R * W (zero = v_0)
    const int ck_stride = p.R * W;
    const size_t base = ((size_t)b * H + row0) * W;
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      on[i] = thread_on && (lr0 + i < n);
      rr[i] = on[i] ? __ldg(reinterpret_cast<const float4*>(p.r + base + (size_t)(lr0 + i) * W + x0))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t sm_base = vi_smem_u32(sm);
    const bool push_up = has_up && thread_on && lr0 == 0;                       // owns strip row 0
    const int dn_i = (has_dn && thread_on && n - 1 >= lr0 && n - 1 < lr0 + NR) ? n - 1 - lr0 : -1;
    // remote byte addresses (tile 0): my top row -> bottom halo (row R+1) of the strip above; my
    // bottom row -> top halo (row 0) of the strip below
    uint32_t up_dst = push_up ? vi_mapa(sm_base, (uint32_t)(cr - 1)) + (uint32_t)(((p.R + 1) * P + 4 + x0) * 4) : 0u;
    uint32_t dn_dst = dn_i >= 0 ? vi_mapa(sm_base, (uint32_t)(cr + 1)) + (uint32_t)((4 + x0) * 4) : 0u;
    uint32_t up_bar0 = push_up ? vi_mapa(vi_smem_u32(&s_mbar[0]), (uint32_t)(cr - 1)) : 0u;
    uint32_t dn_bar0 = dn_i >= 0 ? vi_mapa(vi_smem_u32(&s_mbar[0]), (uint32_t)(cr + 1)) : 0u;
    const uint32_t halo_bytes = (uint32_t)(4 * W) * ((has_up ? 1u : 0u) + (has_dn ? 1u : 0u));
    const bool lane0 = (tid & 31) == 0;
    uint32_t my_tx = tid == 0 ? halo_bytes : 0u;   // thread 0 also announces the halo bytes the neighbours push
    uint32_t P4 = (uint32_t)P * 4u, tile4 = (uint32_t)tile * 4u;
    uint32_t win_a = sm_base + (uint32_t)((lr0 * P + 4 + x0) * 4);   // row above my first cell, tile 0
    uint32_t mbar_a = vi_smem_u32(&s_mbar[0]);
    uint32_t wmax_a = vi_smem_u32(&s_wmax[0][tid >> 5]);
    uint32_t post_a = vi_smem_u32(&s_postbar[0]);
    // keep the loop-invariant addresses in registers: without this the compiler re-derives each of them from the
    // kernel parameters and %cluster_ctaid at every use (a dozen integer instructions per shared-memory access)
    asm volatile("" : "+r"(win_a), "+r"(P4), "+r"(tile4), "+r"(mbar_a), "+r"(wmax_a), "+r"(post_a));
    asm volatile("" : "+r"(up_dst), "+r"(dn_dst), "+r"(up_bar0), "+r"(dn_bar0), "+r"(my_tx));

    // X = r + gamma*v of the own cells into tile `buf` (+ halo pushes), arrive, wait for the phase
    auto exchange = [&](const uint32_t buf, const uint32_t parity) {
      const uint32_t boff = buf ? tile4 : 0u;
      const uint32_t a = win_a + boff;
      float4 X[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        X[i].x = __fadd_rn(rr[i].x, __fmul_rn(v[i].x, gamma));
        X[i].y = __fadd_rn(rr[i].y, __fmul_rn(v[i].y, gamma));
        X[i].z = __fadd_rn(rr[i].z, __fmul_rn(v[i].z, gamma));
        X[i].w = __fadd_rn(rr[i].w, __fmul_rn(v[i].w, gamma));
      }
      if constexpr (PW > 0) {
        vi_static_for<0, NR>([&](auto ic) {
          constexpr int i = decltype(ic)::value;
          if (on[i]) vi_sts_v4<(i + 1) * PW * 4>(a, X[i]);
        });
      } else {
        uint32_t ai = a;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          ai += P4;
          if (on[i]) vi_sts_v4<0>(ai, X[i]);
        }
      }
      if (push_up) vi_st_async_v4(up_dst + boff, X[0], up_bar0 + buf * 8u);
      if (dn_i >= 0) {
        float4 Xd = X[0];
#pragma unroll
        for (int i = 1; i < NR; ++i) if (dn_i == i) Xd = X[i];
        vi_st_async_v4(dn_dst + boff, Xd, dn_bar0 + buf * 8u);
      }
      const uint32_t bar = mbar_a + buf * 8u;
      // the lanes' tile stores are ordered before lane 0's (release) arrival by the warp barrier: one shared-memory
      // atomic per warp per sweep.  expect_tx(0) is a plain arrival, so one instruction serves every warp.
      __syncwarp();
      if (lane0) vi_mbar_arrive_expect_a(bar, my_tx);
      if (!vi_mbar_test_a(bar, parity)) {
        unsigned spins = 0;
        while (!vi_mbar_test_a(bar, parity)) {
          if (++spins > VI2_SPIN_LIMIT) { failed = true; s_fail = 1; break; }
        }
      }
    };

    // one Bellman sweep on tile `buf`: v <- max_a q ; returns max |dv| over the own cells
    auto sweep = [&](const uint32_t buf) -> float {
      const uint32_t ra = win_a + (buf ? tile4 : 0u);
      float a[NR + 2][6];
      if constexpr (PW > 0) {
        vi_static_for<0, NR + 2>([&](auto ic) {
          constexpr int i = decltype(ic)::value;
          const float4 m = vi_lds_v4<i * PW * 4>(ra);
          a[i][0] = vi_lds<i * PW * 4 - 4>(ra); a[i][1] = m.x; a[i][2] = m.y; a[i][3] = m.z; a[i][4] = m.w;
          a[i][5] = vi_lds<i * PW * 4 + 16>(ra);
        });
      } else {
        uint32_t ri = ra;
#pragma unroll
        for (int i = 0; i < NR + 2; ++i) {
          const float4 m = vi_lds_v4<0>(ri);
          a[i][0] = vi_lds<-4>(ri); a[i][1] = m.x; a[i][2] = m.y; a[i][3] = m.z; a[i][4] = m.w; a[i][5] = vi_lds<16>(ri);
          ri += P4;
        }
      }
      float dmax = 0.f;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        float nv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float win[3][3], q[8];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) win[dy][dx] = a[i + dy][j + dx];
          eval_q(win, q);
          nv[j] = fmaxf(vi_max3(q[0], q[1], q[2]), vi_max3(q[3], q[4], vi_max3(q[5], q[6], q[7])));
        }
        // max is exact and order-free, so the three-input form changes no bit
        float d = vi_max3(0.0f, fabsf(__fsub_rn(nv[0], v[i].x)), fabsf(__fsub_rn(nv[1], v[i].y)));
        d = vi_max3(d, fabsf(__fsub_rn(nv[2], v[i].z)), fabsf(__fsub_rn(nv[3], v[i].w)));
        dmax = fmaxf(dmax, on[i] ? d : 0.0f);
        v[i] = make_float4(nv[0], nv[1], nv[2], nv[3]);
      }
      return dmax;
    };

    // warp max of max|dv| in ONE instruction (non-negative floats order like their bit patterns), then a plain
    // store + mbarrier arrive by lane 0: nothing on this path returns a value to wait for
    auto post = [&](const int s, const float dmax) {
      const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(dmax));
      if (lane0) {
        const uint32_t slot = (uint32_t)(s - 1) & (uint32_t)(VI2_RING - 1);
        vi_sts_u32(wmax_a + slot * 128u, wm);
        vi_mbar_arrive_a(post_a + slot * 8u);
      }
    };

    // Blocks of `lag` sweeps (lag even: the tile parity is a compile-time constant inside the block).  At the end of
    // block m (s = m*lag sweeps done) the decisions of every sweep <= J = (m-1)*lag are awaited -- they are at least
    // `lag` sweeps old, so the wait is normally already satisfied -- and only then v_s is snapshotted into slot m & 1:
    // a stop at K in ((m-2)*lag, (m-1)*lag] finds both candidate snapshots, (m-2)*lag and (m-1)*lag, still alive.
    const int lag = p.lag;
    int s = 0;
    uint32_t par = 0;                       // mbarrier phase parity of both tiles for the current pair of sweeps
    for (int m = 1; !failed; ++m) {
      for (int t = 0; t < lag; t += 2) {
        exchange(0u, par);
        float d0 = sweep(0u);
        ++s;
        if (s <= p.max_sweeps) post(s, d0);
        exchange(1u, par);
        float d1 = sweep(1u);
        ++s;
        if (s <= p.max_sweeps) post(s, d1);
        par ^= 1u;
      }
      if (failed) break;
      const int J = min(s - lag, p.max_sweeps);
      if (J >= 1) {
        unsigned spins = 0;
        while (s_done < J) {
          if (s_fail || ++spins > VI2_SPIN_LIMIT) { failed = true; s_fail = 1; break; }
        }
        if (failed) break;
        const int sk = s_stopK;
        if (sk <= J) { K = sk; hit_max = 0; break; }
        if (J == p.max_sweeps) { K = J; hit_max = 1; break; }
      }
      float* dst = ckpt + (m & 1) * ck_stride;      // snapshot v_s, s = m * lag (own cells only)
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (on[i]) *reinterpret_cast<float4*>(dst + (lr0 + i) * W + x0) = v[i];
    }
    int ph = s;                                      // even: the next exchange uses tile 0
    if (!failed) {
      // restore the snapshot at c0 = floor(K / lag) * lag and replay up to sweep K (same arithmetic, same order =>
      // same bits); slot (c0 / lag) & 1 still holds v_c0 (slot 0 starts as zeros = v_0)
      const int c0 = (K / lag) * lag;
      const float* src = ckpt + ((c0 / lag) & 1) * ck_stride;
#pragma unroll
      for (int i = 0; i < NR; ++i)
        if (on[i]) v[i] = *reinterpret_cast<const float4*>(src + (lr0 + i) * W + x0);
      for (int r = c0; r < K && !failed; ++r) {
        exchange((uint32_t)(ph & 1), (uint32_t)((ph >> 1) & 1));
        (void)sweep((uint32_t)(ph & 1));
        ++ph;
      }
    }
    if (!failed) {
      // final pass (vin.py:76-80): q = conv(r + gamma*v_K), pi = softmax_a(q)
      exchange((uint32_t)(ph & 1), (uint32_t)((ph >> 1) & 1));
      const float* X = sm + (ph & 1) * tile;
      const size_t HW = (size_t)H * W;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int lr = lr0 + i;
        if (on[i] && !failed) {
          const int y = row0 + lr;
          if (p.v_out) *reinterpret_cast<float4*>(p.v_out + base + (size_t)lr * W + x0) = v[i];
          if (p.q_out || p.pi_out) {
            const size_t o = (size_t)b * 8 * HW + (size_t)y * W + x0;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              float win[3][3], q[8], e[8];
              const float* c = X + (lr + 1) * P + 4 + x0 + j;
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                win[0][dx + 1] = c[-P + dx];
                win[1][dx + 1] = c[dx];
                win[2][dx + 1] = c[P + dx];
              }
              eval_q(win, q);
              float m = q[0];
#pragma unroll
              for (int k = 1; k < 8; ++k) m = fmaxf(m, q[k]);
              float ssum = 0.0f;
#pragma unroll
              for (int k = 0; k < 8; ++k) { e[k] = expf(__fsub_rn(q[k], m)); ssum = __fadd_rn(ssum, e[k]); }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (p.q_out) p.q_out[o + k * HW + j] = q[k];
                if (p.pi_out) p.pi_out[o + k * HW + j] = __fdiv_rn(e[k], ssum);
              }
            }
          }
        }
      }
    }
    };   // body
    if (my_rows == RT) body(std::integral_constant<int, RT>{});
    else if constexpr (RT > 1) body(std::integral_constant<int, RT - 1>{});
    if (blockIdx.x == 0 && tid == 0 && p.sweeps_out) {
      p.sweeps_out[0] = failed ? -1 : K;
      p.sweeps_out[1] = failed ? -1 : hit_max;
    }
  }
  cluster.sync();   // no CTA may exit while a neighbour can still write into its shared memory
}

// threads an RT <= 2 strip may use: 864 (27 warps, <= 72 registers) when CRESTE_VI_WIDE is set, else 608
static int vi_rt2_limit() { return getenv("CRESTE_VI_WIDE") ? 864 : 608; }

template <int RT, int PW>
static int vi_try_strip(ViStripParams& p, int c, int threads, size_t smem, cudaStream_t st, int* max_clusters_out) {
  auto kern = vi_strip_kernel<RT, PW>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (c > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.B * c);
  cfg.blockDim = dim3(threads + 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  *max_clusters_out = max_clusters;
  if (max_clusters < p.B) {              // every cluster must be co-resident (global delta exchange)
    if (getenv("CRESTE_VI_DEBUG")) {
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, kern);
      fprintf(stderr, "[creste_vi strip] c=%d R=%d RT=%d: max_clusters=%d < B=%d (threads %d, regs %d, maxThreadsPerBlock %d, smem %zu + %zu)\n",
              c, p.R, RT, max_clusters, p.B, threads + 32, fa.numRegs, fa.maxThreadsPerBlock, smem, fa.sharedSizeBytes);
    }
    return 0;
  }
  p.c = c; p.G = p.B * c; p.nt = threads;
  if (getenv("CRESTE_VI_DEBUG"))
    fprintf(stderr, "[creste_vi strip] B=%d H=%d W=%d c=%d R=%d RT=%d threads=%d smem=%zu max_clusters=%d\n",
            p.B, p.H, p.W, c, p.R, RT, threads + 32, smem, max_clusters);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  count_launch();
  return 1;
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
  return 2 * align_up(n * sizeof(float), 256) + 2 * align_up((size_t)(max_sweeps + 64) * 4, 256) + 256;
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
  const size_t gbytes = align_up((size_t)(max_sweeps + 64) * 4, 256);
  w += gbytes;
  p.counter = (unsigned*)(w + gbytes);   // after the second (arrival-count) array
  p.v_out = v_out;
  p.q_out = q_out;
  p.pi_out = pi_out;
  p.sweeps_out = sweeps_out;
  p.B = B; p.H = H; p.W = W;
  p.max_sweeps = max_sweeps;
  p.gamma = gamma; p.thr = thr;
  CRESTE_CUDA(cudaMemsetAsync(p.gdelta, 0, 2 * gbytes + 256, st));
  // ---- strip kernel: largest cluster (16 .. 1 CTAs per sample) that is co-resident for the batch
  if (!getenv("CRESTE_VI_NO_STRIP") && !getenv("CRESTE_VI_NO_CLUSTER") && (W % 4) == 0) {
    ViStripParams sp;
    sp.r = r; sp.v_out = v_out; sp.q_out = q_out; sp.pi_out = pi_out;
    sp.gword = (unsigned long long*)p.gdelta;      // (max, count) pairs over the two arrays
    sp.sweeps_out = sweeps_out;
    sp.B = B; sp.H = H; sp.W = W; sp.max_sweeps = max_sweeps; sp.gamma = gamma; sp.thr = thr;
    const int cgn = W / 4;
    int last_mc = -1;
    for (int c = 16; c >= 1; --c) {
      if (c > H) continue;
      const int R = ceil_div(H, c);
      if ((c - 1) * R >= H) continue;                 // every strip needs at least one row
      // rows per thread RT and row groups rgn: the LAST nfull = R - rgn * (RT - 1) groups own RT rows, the others
      // RT - 1 (mixed only when a warp never spans two groups).  Cost = the largest number of row-sweeps any of the four
      // warp schedulers carries (warp w runs on scheduler w % 4); ties go to the configuration with more warps.
      int RT = 0, rgn = 0, nfull = 0;
      {
        long long best = -1;
        int best_threads = 0;
        const int wpr = cgn >= 32 ? cgn / 32 : 0;        // warps per row group (0: several groups share a warp)
        for (int t = 1; t <= 4; ++t) {
          const int lim = t <= 2 ? vi_rt2_limit() : (t == 3 ? 576 : 512);
          const int g_min = ceil_div(R, t), g_max = (t > 1 && wpr > 0 && cgn % 32 == 0 && !getenv("CRESTE_VI_UNIFORM")) ? R / (t - 1) : g_min;
          for (int g = g_min; g <= g_max; ++g) {
            if ((long long)cgn * g > lim) break;
            const int nf = t > 1 ? R - g * (t - 1) : g;
            if (nf < 0 || nf > g) continue;
            long long load[4] = {0, 0, 0, 0};
            if (wpr > 0) {
              for (int q = 0; q < g; ++q)
                for (int h = 0; h < wpr; ++h) load[(q * wpr + h) & 3] += (q >= g - nf ? t : t - 1);
            } else {
              const int warps = ceil_div(cgn * g, 32);
              for (int w = 0; w < warps; ++w) load[w & 3] += t;
            }
            const long long cost = std::max(std::max(load[0], load[1]), std::max(load[2], load[3]));
            const int threads_g = cgn * g;
            if (best < 0 || cost < best || (cost == best && threads_g > best_threads)) {
              best = cost; best_threads = threads_g; RT = t; rgn = g; nfull = nf;
            }
          }
        }
      }
      if (const char* e = getenv("CRESTE_VI_RT")) {          // experiment knob: force uniform rows-per-thread blocking
        const int t = atoi(e);
        if (t >= 1 && t <= 4 && (long long)cgn * ceil_div(R, t) <= (t <= 2 ? 864 : (t == 3 ? 576 : 512))) {
          RT = t; rgn = ceil_div(R, t); nfull = t > 1 ? R - rgn * (t - 1) : rgn;
          if (nfull < rgn && cgn % 32 != 0) { nfull = rgn; }
        }
      }
      if (!RT) continue;
      int threads = cgn * rgn;
      threads = (threads + 31) / 32 * 32;
      const size_t ssmem = ((size_t)2 * (R + 2) * (W + 8) + (size_t)2 * R * W) * sizeof(float);
      if (ssmem > 200 * 1024) continue;
      sp.rgn = rgn; sp.nfull = nfull;
      sp.R = R;
      sp.lag = (long long)R * W >= 2048 ? 8 : 16;   // even, 2 * lag <= RING
      if (const char* e = getenv("CRESTE_VI_LAG")) { const int l = atoi(e); if (l >= 2 && 2 * l <= VI2_RING && (l & 1) == 0) sp.lag = l; }
      int mc = 0;
      int rc;
#define VI_STRIP_RT(PWV)                                                        \
      rc = RT == 1 ? vi_try_strip<1, PWV>(sp, c, threads, ssmem, st, &mc)       \
         : RT == 2 ? vi_try_strip<2, PWV>(sp, c, threads, ssmem, st, &mc)       \
         : RT == 3 ? vi_try_strip<3, PWV>(sp, c, threads, ssmem, st, &mc)       \
                   : vi_try_strip<4, PWV>(sp, c, threads, ssmem, st, &mc)
      if (W == 256) { VI_STRIP_RT(264); }
      else if (W == 128) { VI_STRIP_RT(136); }
      else if (W == 64) { VI_STRIP_RT(72); }
      else { VI_STRIP_RT(0); }
#undef VI_STRIP_RT
      last_mc = mc;
      if (rc == 1) return 0;
    }
  }
  // ---- cluster-resident path: largest cluster (16, 8, 4, 2, 1 CTAs per sample) that is
  // co-resident for the whole batch and keeps <= 16 cells per thread
  if (!getenv("CRESTE_VI_NO_CLUSTER")) {
    ViClusterParams cp;
    cp.r = r; cp.v_out = v_out; cp.q_out = q_out; cp.pi_out = pi_out;
    cp.gdelta = p.gdelta; cp.garrive = (unsigned*)((char*)p.gdelta + gbytes);
    cp.sweeps_out = sweeps_out;
    cp.B = B; cp.H = H; cp.W = W; cp.max_sweeps = max_sweeps; cp.gamma = gamma; cp.thr = thr;
    for (int c = 16; c >= 1; c >>= 1) {
      if (c > H) continue;
      const int R = ceil_div(H, c);
      if ((c - 1) * R >= H) continue;                 // every strip needs at least one row
      const long long cells = (long long)R * W;
      if (cells > 512LL * 8) continue;    // larger strips run faster on the streamed kernel below
      const size_t csmem = (size_t)2 * (R + 2) * (W + 2) * sizeof(float);
      if (csmem > 200 * 1024) continue;
      cp.R = R;
      const int per = cells <= 512 * 4 ? 4 : 8;
      int threads = (int)((cells + per - 1) / per);
      threads = (threads + 31) / 32 * 32;
      if (threads < 64) threads = 64;
      int rc = per == 4 ? vi_try_cluster<4, 5>(cp, c, threads, csmem, st)
                        : vi_try_cluster<8, 3>(cp, c, threads, csmem, st);
      if (rc == 1) return 0;
    }
  }
  vi_partition(B, H, W, &p.R, &p.G);
  const size_t smem = (size_t)(p.R + 2) * (W + 2) * sizeof(float);
  if (smem > 220 * 1024) {
    set_error("creste_vi_solve: strip of %d rows x %d cols needs %zu B shared memory", p.R, W, smem);
    return CRESTE_ERR_ARG;
  }
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
