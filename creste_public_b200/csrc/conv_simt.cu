// conv_simt.cu -- fp32 implicit-GEMM convolution on CUDA cores (precision mode 0).
//
// The exact-fp32 member of the conv family behind creste_conv2d (include/creste_b200.h): every
// product is a true fp32 FFMA, so this path is the on-device numerical anchor for the tcgen05
// modes, and it also serves the layers whose shapes do not map onto 128-row UMMA tiles (C = 4
// stem, strided convs, K = 1/2/6 projection heads).
//
// GEMM view: M = N*P*Q output pixels, N = K output channels,
// Kdim = R*S*C with k = (r*S + s)*C + c.  NHWC activations make every 4-channel group of the
// A operand one aligned float4 global load; the weight matrix is pre-packed [Kdim][ldw].
// Tile 128 x BN x 16, 256 threads, 8 x TN register micro-tiles, double-buffered shared memory
// with register prefetch (one __syncthreads per k-chunk).  Epilogue fuses folded BatchNorm /
// bias (scale, shift), residual add, ReLU / swish / sigmoid, and writes NHWC or NCHW.
#include <stdlib.h>

#include "common.cuh"

namespace creste {

struct ConvP {
  const float* x; const float* w; const float* scale; const float* shift; const float* gate;
  const float* residual; float* out;
  int N, H, W, C, K, R, S, stride, pad_t, pad_l, P, Q, act, out_nchw;
  int M, Kdim, ldw;
  int vec;      // NHWC output, K % 4 == 0 and 16-byte aligned pointers: 128-bit epilogue loads / stores
  unsigned* amax_out;   // optional: atomicMax of |out| (float bits) -- "amax carried with the tensor" for the
                        // 3xFP16 operand scale of the NEXT tensor-core conv (no separate amax pass over it)
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return v * (1.0f / (1.0f + expf(-v)));   // swish = x * sigmoid(x)
  if (act == 3) return 1.0f / (1.0f + expf(-v));
  return v;
}

constexpr int BM = 128, BK = 16, APITCH = BM + 4;

template <int BN, int TN>
__global__ void __launch_bounds__(256, 2) conv_simt_kernel(ConvP p) {
  constexpr int TM = 8;
  constexpr int NTX = BN / TN;             // threads along N
  static_assert((BM / TM) * NTX == 256, "tile/thread mismatch");
  constexpr int B_F4_PER_THREAD = (BK * BN / 4) / 256 > 0 ? (BK * BN / 4) / 256 : 1;
  constexpr bool B_ALL = (BK * BN / 4) >= 256;
  __shared__ __align__(16) float As[2][BK][APITCH];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  // ---- A gather bookkeeping: this thread loads float4 #kq of rows (t/4) and (t/4 + 64)
  const int kq = t & 3;
  int a_n[2], a_iy0[2], a_ix0[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + (t >> 2) + i * 64;
    a_ok[i] = m < p.M;
    const int mm = a_ok[i] ? m : 0;
    const int ox = mm % p.Q;
    const int oy = (mm / p.Q) % p.P;
    a_n[i] = mm / (p.Q * p.P);
    a_iy0[i] = oy * p.stride - p.pad_t;
    a_ix0[i] = ox * p.stride - p.pad_l;
  }
  // running decode of k = kc*16 + kq*4 -> (r, s, c)
  int k_cur = kq * 4;
  int c_cur = k_cur % p.C;
  int tap = k_cur / p.C;
  int r_cur = tap / p.S, s_cur = tap - r_cur * p.S;

  // ---- B load bookkeeping
  const int b_k = (t * 4) / BN, b_n = (t * 4) % BN;   // for B_F4_PER_THREAD == 1 (BN <= 64)

  float4 a_reg[2];
  float4 b_reg[B_F4_PER_THREAD];

  auto load_a = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int iy = a_iy0[i] + r_cur, ix = a_ix0[i] + s_cur;
      if (a_ok[i] && k_cur < p.Kdim && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
        v = __ldg(reinterpret_cast<const float4*>(
            p.x + (((size_t)a_n[i] * p.H + iy) * p.W + ix) * p.C + c_cur));
        if (p.gate) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)a_n[i] * p.C + c_cur));
          v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
        }
      }
      a_reg[i] = v;
    }
  };
  auto advance_a = [&]() {
    k_cur += BK;
    c_cur += BK;
    while (c_cur >= p.C) {
      c_cur -= p.C;
      if (++s_cur == p.S) { s_cur = 0; ++r_cur; }
    }
  };
  auto load_b = [&](int kc) {
    if (B_ALL) {
#pragma unroll
      for (int j = 0; j < B_F4_PER_THREAD; ++j) {
        const int f = t + j * 256;
        const int kk = (f * 4) / BN, nn = (f * 4) % BN;
        const int k = kc * BK + kk, n = n0 + nn;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < p.Kdim && n < p.ldw) v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)k * p.ldw + n));
        b_reg[j] = v;
      }
    } else {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < BK * BN / 4) {
        const int k = kc * BK + b_k, n = n0 + b_n;
        if (k < p.Kdim && n < p.ldw) v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)k * p.ldw + n));
      }
      b_reg[0] = v;
    }
  };
  auto store_ab = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = (t >> 2) + i * 64;
      As[buf][kq * 4 + 0][m] = a_reg[i].x;
      As[buf][kq * 4 + 1][m] = a_reg[i].y;
      As[buf][kq * 4 + 2][m] = a_reg[i].z;
      As[buf][kq * 4 + 3][m] = a_reg[i].w;
    }
    if (B_ALL) {
#pragma unroll
      for (int j = 0; j < B_F4_PER_THREAD; ++j) {
        const int f = t + j * 256;
        *reinterpret_cast<float4*>(&Bs[buf][(f * 4) / BN][(f * 4) % BN]) = b_reg[j];
      }
    } else if (t < BK * BN / 4) {
      *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = b_reg[0];
    }
  };

  const int tx = t % NTX, ty = t / NTX;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  const int nchunks = ceil_div(p.Kdim, BK);
  load_a();
  load_b(0);
  store_ab(0);
  __syncthreads();
  for (int kc = 0; kc < nchunks; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nchunks) {
      advance_a();
      load_a();
      load_b(kc + 1);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if constexpr (TN == 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][BN / 2 + tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
      } else if constexpr (TN == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else {
        const float2 b0 = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        b[0] = b0.x; b[1] = b0.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kc + 1 < nchunks) {
      store_ab(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue
  float amx = 0.0f;
  auto post_amax = [&]() {
    if (p.amax_out) {
      amx = warp_max(amx);
      if ((t & 31) == 0 && amx > 0.0f) atomicMax(p.amax_out, __float_as_uint(amx));
    }
  };
  // vector path: a thread's output columns come in groups of 4 (2 for TN == 2) contiguous channels, so the
  // folded BN factors, the residual and the store are one 128-bit access per group -- 16 lanes write 256
  // contiguous bytes of a pixel row.  (The scalar form below wrote 4 bytes per lane at a 16-byte stride: a
  // quarter of every 32-byte sector, and 4x the store instructions; it was 23 % of the forward.)
  if (p.vec) {
    constexpr int G = TN >= 4 ? 4 : 2;            // channels per group
#pragma unroll
    for (int jg = 0; jg < TN / G; ++jg) {
      const int n = n0 + ((TN == 8) ? (jg ? BN / 2 + tx * 4 : tx * 4) : tx * TN);
      if (n >= p.K) continue;
      float sc[G], sh[G];
      if (G == 4) {
        const float4 s4 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 h4 = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        sc[0] = s4.x; sc[1] = s4.y; sc[2 % G] = s4.z; sc[3 % G] = s4.w;
        sh[0] = h4.x; sh[1] = h4.y; sh[2 % G] = h4.z; sh[3 % G] = h4.w;
      } else {
        const float2 s2 = p.scale ? __ldg(reinterpret_cast<const float2*>(p.scale + n)) : make_float2(1.f, 1.f);
        const float2 h2 = p.shift ? __ldg(reinterpret_cast<const float2*>(p.shift + n)) : make_float2(0.f, 0.f);
        sc[0] = s2.x; sc[1] = s2.y; sh[0] = h2.x; sh[1] = h2.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        float v[G];
#pragma unroll
        for (int u = 0; u < G; ++u) v[u] = fmaf(acc[i][jg * G + u], sc[u], sh[u]);
        float* dst = p.out + (size_t)m * p.K + n;
        if (G == 4) {
          if (p.residual) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)m * p.K + n));
            v[0] += r4.x; v[1] += r4.y; v[2 % G] += r4.z; v[3 % G] += r4.w;
          }
          const float4 o4 = make_float4(apply_act(v[0], p.act), apply_act(v[1], p.act),
                                        apply_act(v[2 % G], p.act), apply_act(v[3 % G], p.act));
          amx = fmaxf(fmaxf(amx, fmaxf(fabsf(o4.x), fabsf(o4.y))), fmaxf(fabsf(o4.z), fabsf(o4.w)));
          *reinterpret_cast<float4*>(dst) = o4;
        } else {
          if (p.residual) {
            const float2 r2 = __ldg(reinterpret_cast<const float2*>(p.residual + (size_t)m * p.K + n));
            v[0] += r2.x; v[1] += r2.y;
          }
          const float2 o2 = make_float2(apply_act(v[0], p.act), apply_act(v[1], p.act));
          amx = fmaxf(amx, fmaxf(fabsf(o2.x), fabsf(o2.y)));
          *reinterpret_cast<float2*>(dst) = o2;
        }
      }
    }
    post_amax();
    return;
  }
  const int PQ = p.P * p.Q;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + ((TN == 8) ? ((j >= 4) ? (BN / 2 + tx * 4 + j - 4) : (tx * 4 + j)) : (tx * TN + j));
    if (n >= p.K) continue;
    const float sc = p.scale ? __ldg(p.scale + n) : 1.0f;
    const float sh = p.shift ? __ldg(p.shift + n) : 0.0f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int m = m0 + ty * TM + i;
      if (m >= p.M) continue;
      float v = fmaf(acc[i][j], sc, sh);
      if (p.residual) v += __ldg(p.residual + (size_t)m * p.K + n);
      v = apply_act(v, p.act);
      amx = fmaxf(amx, fabsf(v));
      if (p.out_nchw) {
        const int img = m / PQ, pix = m - img * PQ;
        p.out[((size_t)img * p.K + n) * PQ + pix] = v;
      } else {
        p.out[(size_t)m * p.K + n] = v;
      }
    }
  }
  post_amax();
}

int conv_simt_launch(const creste_conv_desc* d, const float* x, const float* w, int ldw,
                     const float* scale, const float* shift, const float* gate,
                     const float* residual, float* out, unsigned* amax_out, cudaStream_t st) {
  ConvP p;
  p.amax_out = amax_out;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.gate = gate; p.residual = residual;
  p.out = out;
  p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C; p.K = d->K; p.R = d->R; p.S = d->S;
  p.stride = d->stride; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.P = d->P; p.Q = d->Q;
  p.act = d->act; p.out_nchw = d->out_nchw;
  p.M = d->N * d->P * d->Q;
  p.Kdim = d->R * d->S * d->C;
  p.ldw = ldw;
  auto al16 = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15u) == 0; };
  p.vec = (!d->out_nchw && d->K % 4 == 0 && al16(out) && al16(residual) && al16(scale) && al16(shift)) ? 1 : 0;
  if (getenv("CRESTE_SIMT_SCALAR_EPILOGUE")) p.vec = 0;
  const int gm = ceil_div(p.M, BM);
  if (d->K > 64) {
    dim3 grid(gm, ceil_div(d->K, 128));
    conv_simt_kernel<128, 8><<<grid, 256, 0, st>>>(p);
  } else if (d->K > 32) {
    dim3 grid(gm, 1);
    conv_simt_kernel<64, 4><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(gm, 1);
    conv_simt_kernel<32, 2><<<grid, 256, 0, st>>>(p);
  }
  return launch_check("conv_simt_kernel");
}

}  // namespace creste
