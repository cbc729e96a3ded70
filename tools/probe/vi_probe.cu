// vi_probe.cu -- two hardware questions behind the value-iteration kernel (profiles/r2_vi_experiments.md):
//  (1) what does a packed FFMA2 (fma.rn.f32x2) cost in issue slots / pipe time next to two scalar FFMAs,
//  (2) how many thread-block clusters of size c are REALLY co-resident on this B200 (the occupancy query vs a launch
//      whose CTAs wait for each other with a time-out).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/vi_probe.bin tools/probe/vi_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ unsigned long long fma2_imm(unsigned long long a, unsigned long long c) {
  unsigned long long d, k = 0x3dcccccd3dcccccdULL;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(k), "l"(c)); return d;
}

template <int MODE>   // 0: scalar FFMA imm, 16 chains; 1: FFMA2, 8 packed chains (16 FMAs); 2: scalar + FMNMX mix; 3: FFMA2 + FMNMX mix
__global__ void fma_probe(float* out, long long* cyc, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
  float mx = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (MODE == 0 || MODE == 2) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], 0.1f, x[(i + 1) & 15]);
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) mx = fmaxf(mx, fmaxf(x[i], x[i + 1]));
      }
    }
  } else {
    unsigned long long p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = pack2(x[2 * i], x[2 * i + 1]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2_imm(p[i], p[(i + 1) & 7]);
      if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i]));
          mx = fmaxf(mx, fmaxf(a, b));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) asm("mov.b64 {%0,%1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(p[i]));
  }
  long long t1 = clock64();
  float s = mx;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void cluster_probe(unsigned* counter, int* smid_out, int* ok_out, unsigned total, long long budget_cycles) {
  extern __shared__ float dyn[];
  unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  if (threadIdx.x == 0) {
    dyn[0] = 0.f;
    smid_out[blockIdx.x] = (int)sm;
    __threadfence();
    atomicAdd(counter, 1u);
    long long t0 = clock64();
    int ok = 0;
    while (clock64() - t0 < budget_cycles) {
      if (*((volatile unsigned*)counter) >= total) { ok = 1; break; }
    }
    ok_out[blockIdx.x] = ok;
  }
  __syncthreads();
  cg::this_cluster().sync();
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  // ---- (1) FFMA2
  float* out; long long* cyc; CK(cudaMalloc(&out, 1024 * 148 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
  const int iters = 4096;
  for (int threads : {128, 256, 512, 1024}) {
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) fma_probe<0><<<1, threads>>>(out, cyc, iters);
        if (mode == 1) fma_probe<1><<<1, threads>>>(out, cyc, iters);
        if (mode == 2) fma_probe<2><<<1, threads>>>(out, cyc, iters);
        if (mode == 3) fma_probe<3><<<1, threads>>>(out, cyc, iters);
        CK(cudaDeviceSynchronize());
      }
      long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
      double fmas = 16.0 * iters * threads;
      printf("fma_probe mode %d (%s) threads %4d: %lld cycles, %.1f FMA/clk/SM\n", mode,
             mode == 0 ? "FFMA imm" : mode == 1 ? "FFMA2 imm" : mode == 2 ? "FFMA+FMNMX" : "FFMA2+FMNMX", threads, c, fmas / (double)c);
    }
  }
  // ---- (2) clusters
  unsigned* counter; int *smid, *ok; CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&smid, 4096)); CK(cudaMalloc(&ok, 4096));
  CK(cudaFuncSetAttribute(cluster_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  CK(cudaFuncSetAttribute(cluster_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int threads : {416, 608}) for (size_t smem : {(size_t)1024, (size_t)120 * 1024}) for (int c = 8; c <= 16; ++c) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1; cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cfg.gridDim = dim3(c * 8);
    int q = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&q, cluster_probe, &cfg);
    if (e != cudaSuccess) { printf("c=%d query error %s\n", c, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    int best = 0;
    std::vector<int> hs(4096 / 4);
    for (int n = std::max(1, q - 1); n <= 148 / c; ++n) {
      CK(cudaMemset(counter, 0, 4)); CK(cudaMemset(ok, 0, 4096));
      cfg.gridDim = dim3(c * n);
      unsigned total = c * n; long long budget = 40000000LL;   // ~20 ms
      e = cudaLaunchKernelEx(&cfg, cluster_probe, counter, smid, ok, total, budget);
      if (e != cudaSuccess) { printf("c=%d n=%d launch error %s\n", c, n, cudaGetErrorString(e)); cudaGetLastError(); break; }
      CK(cudaDeviceSynchronize());
      std::vector<int> ho(c * n); CK(cudaMemcpy(ho.data(), ok, 4 * c * n, cudaMemcpyDeviceToHost));
      bool all = true; for (int v : ho) all = all && v;
      if (all) { best = n; CK(cudaMemcpy(hs.data(), smid, 4 * c * n, cudaMemcpyDeviceToHost)); } else break;
    }
    printf("threads %d smem %zu cluster %2d: occupancy query %d, co-resident by launch %d (%d SMs)\n", threads, smem, c, q, best, best * c);
    if (threads == 608 && smem > 4096 && best > 0) {
      for (int k = 0; k < best; ++k) {
        std::vector<int> v(hs.begin() + k * c, hs.begin() + (k + 1) * c); std::sort(v.begin(), v.end());
        printf("   cluster %d smids:", k); for (int s : v) printf(" %d", s); printf("\n");
      }
    }
  }
  return 0;
}
