// conv1x1.cu -- streaming 1x1 convolution for short reductions (exact fp32, precision mode 0).
//
// The MBConv expand / project convs of the EfficientNet trunk (reference creste/models/blocks/effnet.py:8-98 through
// efficientnet_pytorch's MBConvBlock) at the 256x480 .. 64x120 resolutions have C <= 144 input channels and move
// 100-400 MB each: they are HBM-bound, and the generic implicit-GEMM tile of conv_simt.cu (one 128-pixel tile per
// CTA, a single k-chunk, index decode + two barriers per 56 KB of traffic) ran them at 0.2-0.3 of the HBM rate
// (profiles/r2_launches_3xfp16.md: 363 us for 16->96 @256x480 x 8 frames = 440 MB).  Here:
//   * the whole weight matrix [C][K] sits in shared memory for the lifetime of a persistent CTA;
//   * for a 1x1 / stride-1 conv a run of pixels is ONE contiguous block of memory on both sides (NHWC), so the x tile
//     is a flat coalesced float4 copy and the output tile is written as consecutive float4s: thread t owns the
//     4-channel group kg = t % (K/4) of the pixels s + S*j (s = t / (K/4)), i.e. for every j the CTA stores
//     blockDim.x consecutive float4s;
//   * a thread accumulates 4 channels x PP pixels in registers (4 + PP shared loads per 16*PP FFMA).
// Arithmetic: the same in-order fp32 FMA chain over c = 0 .. C-1 from acc = 0 and the same epilogue
// (fmaf(acc, scale, shift), + residual, activation) as conv_simt_kernel, so the results are bit-identical to it.
#include <stdlib.h>

#include "common.cuh"

namespace creste {

struct Conv1x1P {
  const float* x; const float* w; const float* scale; const float* shift; const float* gate;
  const float* residual; float* out; unsigned* amax_out;
  long long M, HW;
  int C, K, ldw, act, S, CP;
};

__device__ __forceinline__ float c11_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return v * (1.0f / (1.0f + expf(-v)));   // swish = x * sigmoid(x)
  if (act == 3) return 1.0f / (1.0f + expf(-v));
  return v;
}

template <int PP>
__global__ void __launch_bounds__(256) conv1x1_stream_kernel(Conv1x1P p) {
  extern __shared__ __align__(16) float c11_smem[];
  float* ws = c11_smem;                       // [C][K]
  float* xs = c11_smem + (size_t)p.C * p.K;    // [S * PP][CP]
  const int t = threadIdx.x, T = blockDim.x;
  const int KG = p.K >> 2, C4 = p.C >> 2;
  const int kg = t % KG, s = t / KG;
  const int tpix = p.S * PP;

  for (int i = t; i < p.C * KG; i += T) {
    const int c = i / KG, k4 = i - c * KG;
    *reinterpret_cast<float4*>(ws + (size_t)c * p.K + k4 * 4) =
        __ldg(reinterpret_cast<const float4*>(p.w + (size_t)c * p.ldw + k4 * 4));
  }
  const float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + kg * 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 sh = p.shift ? __ldg(reinterpret_cast<const float4*>(p.shift + kg * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
  float amx = 0.0f;

  for (long long p0 = (long long)blockIdx.x * tpix; p0 < p.M; p0 += (long long)gridDim.x * tpix) {
    __syncthreads();                          // previous tile fully consumed (and the weights visible)
    const long long left = p.M - p0;
    const int np = left < tpix ? (int)left : tpix;
    const float4* src = reinterpret_cast<const float4*>(p.x + p0 * p.C);
    // division-free walk over the tile's float4s: f = t, t + T, ... <-> (pix, c4) advances by (T / C4, T % C4);
    // the image index of a pixel (for the SE gate) follows the same way from the tile's first pixel
    const long long img0 = p0 / p.HW;
    const int rem0 = (int)(p0 - img0 * p.HW);                 // first pixel's offset inside its image
    int pix = t / C4, c4 = t - pix * C4;
    const int dpix = T / C4, dc4 = T - dpix * C4;
    for (int f = t; f < np * C4; f += T) {
      float4 v = __ldg(src + f);
      if (p.gate) {
        long long img = img0;
        for (long long r = (long long)rem0 + pix; r >= p.HW; r -= p.HW) ++img;
        const float4 g = __ldg(reinterpret_cast<const float4*>(p.gate + img * p.C + c4 * 4));
        v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
      }
      *reinterpret_cast<float4*>(xs + (size_t)pix * p.CP + c4 * 4) = v;
      pix += dpix; c4 += dc4;
      if (c4 >= C4) { c4 -= C4; ++pix; }
    }
    __syncthreads();
    float4 acc[PP];
#pragma unroll
    for (int j = 0; j < PP; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wp = ws + kg * 4;
    const float* xp = xs + (size_t)s * p.CP;
    const int xstep = p.S * p.CP;
    for (int c = 0; c < p.C; c += 4) {
      const float4 w0 = *reinterpret_cast<const float4*>(wp + (size_t)(c + 0) * p.K);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + (size_t)(c + 1) * p.K);
      const float4 w2 = *reinterpret_cast<const float4*>(wp + (size_t)(c + 2) * p.K);
      const float4 w3 = *reinterpret_cast<const float4*>(wp + (size_t)(c + 3) * p.K);
#pragma unroll
      for (int j = 0; j < PP; ++j) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + (size_t)j * xstep + c);
        acc[j].x = fmaf(xv.x, w0.x, acc[j].x); acc[j].y = fmaf(xv.x, w0.y, acc[j].y);
        acc[j].z = fmaf(xv.x, w0.z, acc[j].z); acc[j].w = fmaf(xv.x, w0.w, acc[j].w);
        acc[j].x = fmaf(xv.y, w1.x, acc[j].x); acc[j].y = fmaf(xv.y, w1.y, acc[j].y);
        acc[j].z = fmaf(xv.y, w1.z, acc[j].z); acc[j].w = fmaf(xv.y, w1.w, acc[j].w);
        acc[j].x = fmaf(xv.z, w2.x, acc[j].x); acc[j].y = fmaf(xv.z, w2.y, acc[j].y);
        acc[j].z = fmaf(xv.z, w2.z, acc[j].z); acc[j].w = fmaf(xv.z, w2.w, acc[j].w);
        acc[j].x = fmaf(xv.w, w3.x, acc[j].x); acc[j].y = fmaf(xv.w, w3.y, acc[j].y);
        acc[j].z = fmaf(xv.w, w3.z, acc[j].z); acc[j].w = fmaf(xv.w, w3.w, acc[j].w);
      }
    }
    if (s < p.S) {
#pragma unroll
      for (int j = 0; j < PP; ++j) {
        const int pix = s + p.S * j;
        if (pix >= np) continue;
        const size_t o = (size_t)(p0 + pix) * p.K + kg * 4;
        float4 v;
        v.x = fmaf(acc[j].x, sc.x, sh.x); v.y = fmaf(acc[j].y, sc.y, sh.y);
        v.z = fmaf(acc[j].z, sc.z, sh.z); v.w = fmaf(acc[j].w, sc.w, sh.w);
        if (p.residual) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + o));
          v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
        }
        v = make_float4(c11_act(v.x, p.act), c11_act(v.y, p.act), c11_act(v.z, p.act), c11_act(v.w, p.act));
        amx = fmaxf(fmaxf(amx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        *reinterpret_cast<float4*>(p.out + o) = v;
      }
    }
  }
  if (p.amax_out) {
    amx = warp_max(amx);
    if ((t & 31) == 0 && amx > 0.0f) atomicMax(p.amax_out, __float_as_uint(amx));
  }
}

// 1: launched; 0: shape not served (the caller falls back to conv_simt_kernel); < 0: error
int conv1x1_stream_launch(const creste_conv_desc* d, const float* x, const float* w, int ldw, const float* scale,
                          const float* shift, const float* gate, const float* residual, float* out,
                          unsigned* amax_out, cudaStream_t st) {
  if (getenv("CRESTE_NO_CONV1X1")) return 0;
  if (d->R != 1 || d->S != 1 || d->stride != 1 || d->pad_t != 0 || d->pad_l != 0 || d->out_nchw) return 0;
  if (d->P != d->H || d->Q != d->W) return 0;
  // C <= 96 for any K <= 256 (measured: wider reductions are no faster here than conv_simt_kernel); round 2d: also the
  // PROJECT shape -- up to 192 input channels onto <= 32 output channels (MBConv 144 -> 24; 144 -> 40 measured no faster; and the data
  // gradient of the expand convs), which the generic kernel ran at 1/6 of the HBM rate with its 32-wide N tile
  const bool project = d->C > 96 && d->C <= 192 && d->K <= 32 && !getenv("CRESTE_C11_NO_PROJECT");
  if (d->K % 4 != 0 || d->K > 256 || d->C % 4 != 0 || (d->C > 96 && !project)) return 0;
  auto al16 = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15u) == 0; };
  if (!al16(x) || !al16(w) || !al16(out) || !al16(residual) || !al16(scale) || !al16(shift) || !al16(gate) || ldw % 4 != 0)
    return 0;
  const int KG = d->K / 4;
  const int CP = d->C + 4;                                  // row pitch: neighbouring pixels land 4 banks apart
  const size_t wbytes = (size_t)d->C * d->K * sizeof(float);
  if (wbytes > 64 * 1024) return 0;
  // pixels per step S (threads T = S * KG in [128, 256]) and pixels per thread PP: the largest tile whose x rows fit
  // the shared-memory budget; few output channels mean many pixels per step, so S is halved until the tile fits
  const size_t budget = (project ? 56 : 40) * 1024;
  const int pp_max = getenv("CRESTE_C11_PP") ? atoi(getenv("CRESTE_C11_PP")) : 4;   // 4: <= 64 registers, 4 CTAs per SM (measured faster than 8)
  int S = 0, T = 0, PP = 0;
  for (int s_try = 256 / KG; s_try >= 1 && s_try * KG >= 128 && !PP; s_try = s_try / 2) {
    for (int pp : {8, 4, 2})
      if (pp <= pp_max && (size_t)s_try * pp * CP * sizeof(float) <= budget) { PP = pp; S = s_try; T = s_try * KG; break; }
  }
  if (!PP) return 0;
  const size_t smem = wbytes + (size_t)S * PP * CP * sizeof(float);
  Conv1x1P p;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.gate = gate; p.residual = residual; p.out = out;
  p.amax_out = amax_out;
  p.M = (long long)d->N * d->H * d->W; p.HW = (long long)d->H * d->W;
  p.C = d->C; p.K = d->K; p.ldw = ldw; p.act = d->act; p.S = S; p.CP = CP;
  const long long tiles = (p.M + (long long)S * PP - 1) / ((long long)S * PP);
  const int sms = num_sms() > 0 ? num_sms() : 148;
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)(tiles < (long long)sms * per_sm ? tiles : (long long)sms * per_sm);
#define CRESTE_C11(PPV)                                                                                   \
  {                                                                                                       \
    static bool attr_done = false;                                                                        \
    if (!attr_done) {                                                                                     \
      if (cudaFuncSetAttribute(conv1x1_stream_kernel<PPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                               112 * 1024) != cudaSuccess) { cudaGetLastError(); return 0; }              \
      attr_done = true;                                                                                   \
    }                                                                                                     \
    conv1x1_stream_kernel<PPV><<<grid, T, smem, st>>>(p);                                                 \
  }
  if (PP == 8) CRESTE_C11(8)
  else if (PP == 4) CRESTE_C11(4)
  else CRESTE_C11(2)
#undef CRESTE_C11
  const int rc = launch_check("conv1x1_stream_kernel");
  return rc == 0 ? 1 : rc;
}

}  // namespace creste
