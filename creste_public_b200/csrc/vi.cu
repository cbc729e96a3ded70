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
  return 2 * align_up(n * sizeof(float), 256) + 2 * align_up((size_t)(max_sweeps + 1) * 4, 256) + 256;
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
  p.counter = (unsigned*)(w + gbytes);   // after the second (arrival-count) array
  p.v_out = v_out;
  p.q_out = q_out;
  p.pi_out = pi_out;
  p.sweeps_out = sweeps_out;
  p.B = B; p.H = H; p.W = W;
  p.max_sweeps = max_sweeps;
  p.gamma = gamma; p.thr = thr;
  CRESTE_CUDA(cudaMemsetAsync(p.gdelta, 0, 2 * gbytes + 256, st));
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
