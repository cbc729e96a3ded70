// conv_tc.cu -- implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA (precision modes 1 = 3xTF32 split, 2 = single-pass TF32).
//
// Covers the stride-1 1x1 / 3x3 convolutions that hold 97 % of the frame's flops (SURVEY.md
// App. D1: effnet up1-3, eff.conv, depth head, dino head, fusion, ResNet-18 BEV layers, the three
// DeconvHeads).  GEMM view: M = output pixels, N = output channels, K = R*S*C.
//
// Data movement
//   A (activations, NHWC fp32): one 4-D tensor map (C, W, H, N).  An M-tile is a spatial box of
//     Wbox x Hbox = 128 output pixels; for filter tap (r,s) and channel block cb the producer
//     issues ONE cp.async.bulk.tensor.4d with box {32 ch, Wbox, Hbox, 1} at coordinates
//     (32*cb, x0+s-pad, y0+r-pad, n).  TMA zero-fills everything outside the tensor, which is the
//     conv's zero padding, the ragged right/bottom tiles and the C % 32 tail in one mechanism.
//     The box lands as 128 rows x 128 B with the 128-byte swizzle = the K-major UMMA operand layout.
//   B (weights): packed [Npad][R*S*Cpad] (Cpad = C rounded up to 32, zero filled), 2-D tensor map,
//     box {32, BLOCK_N}; same swizzle.
//   D: fp32 accumulator in TMEM (128 lanes x BLOCK_N columns), read back with tcgen05.ld by four
//     epilogue warps (one TMEM lane quarter each) that apply folded BN / bias, residual, activation
//     and store NHWC.
//
// Precision: tf32 MMAs read fp32 words and ignore the low 13 mantissa bits.  Mode 1 feeds
// pre-rounded hi = rna_tf32(x) and lo = rna_tf32(x - hi) operands (the split pre-pass
// `tf32_split_kernel`, which also applies the optional SE gate) and issues
// D += Alo*Bhi + Ahi*Blo + Ahi*Bhi per k-step: products carry ~21 mantissa bits, accumulation is
// fp32, which measures as fp32-faithful on the costmap (DESIGN.md "Precision").  Round-to-nearest
// in the split matters: truncation splits bias every product the same way and cost 10x accuracy.
//
// CTA pairs: two CTAs on adjacent M tiles run as one cta_group::2 pair (M = 256 MMAs issued by the
// leader, each CTA staging its own A tile and half of the weight tile), which halves the
// shared-memory operand traffic per MMA -- the single-CTA form measured ~50 % of the tensor peak.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer (one elected
// lane issues tcgen05.mma; tcgen05.commit releases smem stages / signals the epilogue),
// warps 2-9 = epilogue.  mbarrier ring of NSTAGES smem stages.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace creste {

// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                               int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// ---- cta_group::2 (CTA pair) variants: TMA signals the LEADER CTA's mbarrier (peer bit cleared)
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// 2-SM load multicast to the CTAs in `mask`: the tile lands at the same offset in every destination
// CTA and the transaction bytes are signalled on each destination's pair-leader barrier
__device__ __forceinline__ void tma_load_2d_2sm_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                                   int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(smem_u32(dst)),
      "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_BIT_MASK), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (UMMA::SmemDescriptor):
// start>>4 | LBO(ignored)=1 | SBO = 1024 B (8 rows x 128 B) | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, K-major A and B, M = 128
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n, int m = 128) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// UMMA::InstrDescriptor for kind::f16 with fp16 A/B (format 0), fp32 accumulate, K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int n, int m = 128) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ unsigned long long* g_tc_dbg = nullptr;      // optional per-CTA timeline (tools/conv_timeline.py)
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(slot)                                                                  \
  do {                                                                                  \
    if (g_tc_dbg && threadIdx.x == (slot == 3 || slot == 4 ? 64 : (slot == 2 ? 32 : 0))) \
      g_tc_dbg[(size_t)blockIdx.x * 8 + slot] = gtimer();                               \
  } while (0)

// ------------------------------------------------------------------------------------ kernels
struct TcParams {
  const float* scale; const float* shift; const float* residual; float* out;
  int N, P, Q, K;              // output NHWC [N,P,Q,K]
  int R, S, pad_t, pad_l, stride;
  int cblocks;                 // ceil(C / 32)
  int wbox, hbox, tiles_x, tiles_y;
  int block_n, act, split;     // split: 1 = 3xTF32 (hi/lo operands), 0 = single TF32
  int out_nchw;
  int cl;                      // cluster size along the M tiles; the B tile is TMA-multicast in cl slices
  // 3xFP16 mode: operands are fp16 hi / lo of (x * s_a) and (w * s_w[k]) with power-of-two scales;
  // the epilogue undoes them: (main + cross * cross_scale) * act_inv[0] * w_inv[k]
  const float* act_inv; const float* w_inv; float cross_scale;
  int kelems;                  // operand elements per 128-byte swizzle row: 32 (tf32) or 64 (fp16)
  int n_tiles;                 // number of N tiles (output-channel blocks)
  unsigned* amax_out;          // optional: atomicMax of |out| (float bits), carried with the output tensor
  // optional (3xFP16 / fp16 modes): the output written as the NEXT tensor-core conv's operand -- fp16 hi / lo of
  // out * s_out with a power-of-two s_out derived from an a-priori bound of max|out|:
  //   |out| <= max|x| * max_k(sum|w_k| * |scale_k|) + max_k|shift_k| = (2^15 / s_in) * bound_mul + bound_add
  // (any power-of-two scale with amax * s in [2^-4, 2^15] keeps the split at 22 bits, so the loose bound costs
  // nothing).  `out` may then be null: the fp32 tensor is not written and the consumer's split pre-pass disappears.
  uint2* out_hi; uint2* out_lo; float* out_scal;
  float bound_mul, bound_add;
  // max|x| for that bound: the TRUE maximum carried with the input (device float[1]) when there is one -- chained
  // a-priori bounds would compound (each is 2^7 .. 2^10 loose) -- else 2^15 / s_in from the operand's scale record
  const float* in_amax;
};

// optional extra outputs of one conv launch (see TcParams::out_hi)
struct TcSplitOut {
  void* hi; void* lo; float* scal; float bound_mul, bound_add; const float* in_amax;
};

constexpr int TC_THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int A_TILE_BYTES = 128 * 128;   // 128 rows x 32 fp32

__device__ __forceinline__ float tc_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return v / (1.0f + expf(-v));
  if (act == 3) return 1.0f / (1.0f + expf(-v));
  return v;
}

// PAIR = true: the two CTAs of a cluster form a tcgen05 CTA pair (cta_group::2).  Each CTA stages
// its own 128-pixel A tile and HALF of the weight tile (block_n/2 rows); the leader CTA issues
// M = 256 MMAs that read both CTAs' shared memory and write each CTA's half of D into that CTA's
// TMEM.  Operand bytes read from shared memory per MMA-cycle halve -- the single-CTA form is bound
// by shared-memory bandwidth ((128 + 256) rows x 32 B per 128-cycle MMA + the TMA fill).
template <bool PAIR, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 2)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               TcParams p, int nstages) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_mul[256], s_add[256];      // per-output-channel epilogue factors of this N tile
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nops = p.split ? 2 : 1;
  TC_STAMP(0);
  // cluster of p.cl CTAs = p.cl / 2 CTA pairs on consecutive M tiles, all on the same N tile
  const uint32_t crank = PAIR ? cluster_ctarank() : 0;
  const uint32_t rank = crank & 1u;                        // rank inside the CTA pair
  const bool leader = rank == 0;
  const int npairs = PAIR ? p.cl / 2 : 1;
  const uint16_t pair_mask = (uint16_t)(3u << (crank & ~1u));          // this pair's two CTAs
  const uint16_t all_mask = (uint16_t)((1u << p.cl) - 1u);
  uint16_t bcast_mask = 0;                                 // CTAs that stage the same weight half
  for (int j = 0; j < npairs; ++j) bcast_mask |= (uint16_t)(1u << (rank + 2 * j));
  const int b_rows = PAIR ? p.block_n / 2 : p.block_n;     // weight rows staged by THIS CTA
  const int b_tile_bytes = b_rows * 128;
  const int stage_bytes = nops * (A_TILE_BYTES + b_tile_bytes);

  // tile coordinates.  The N tiles of one M tile (pair) are adjacent in launch order, so the
  // activation tile they share is read from DRAM once and from L2 afterwards (measured: with the
  // N tile on blockIdx.y the 487 MB up3 input was fetched twice).
  const int cls = PAIR ? p.cl : 1;
  const int grp = blockIdx.x / cls;                       // (m tile group, n tile) in launch order
  const int n_tile = grp % p.n_tiles;
  int t = (grp / p.n_tiles) * cls + (int)(blockIdx.x % cls);
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int img = t / p.tiles_y;
  const int x0 = tx * p.wbox, y0 = ty * p.hbox;
  const int n0 = n_tile * p.block_n;
  const int kblocks = p.R * p.S * p.cblocks;
  // split mode keeps TWO accumulators: hi*hi in columns [0, block_n) and the two small cross
  // terms in [acc2, acc2 + block_n).  The tensor core truncates (round-toward-zero) the fp32
  // accumulator at every MMA, an error proportional to |acc| per accumulation; separating the
  // cross terms (2^-11 of the main sum) means the large accumulator is rounded K/8 times
  // instead of 3K/8 times.  The epilogue adds the two in fp32 round-to-nearest.
  const uint32_t acc_cols = p.block_n <= 32 ? 32 : (p.block_n <= 64 ? 64 : (p.block_n <= 128 ? 128 : 256));
  const uint32_t acc2 = acc_cols;
  const uint32_t tmem_cols = p.split ? 2 * acc_cols : acc_cols;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a_hi);
    prefetch_tmap(&map_b_hi);
    if (p.split) { prefetch_tmap(&map_a_lo); prefetch_tmap(&map_b_lo); }
    // a stage is free again once EVERY pair of the cluster has consumed it (its weight slices are
    // written by the other pairs' multicast loads)
    for (int i = 0; i < nstages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], npairs); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2sm(&tmem_base_smem, tmem_cols);
    else tmem_alloc(&tmem_base_smem, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();       // barrier inits + TMEM allocation of both CTAs in place
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  TC_STAMP(1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs of a pair; completion bytes land on the leader's barrier) =====
    if (elect_one()) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int stage = kb % nstages;
        const uint32_t parity = ((kb / nstages) & 1) ^ 1;
        mbar_wait(&empty_bar[stage], parity);
        const int cb = kb % p.cblocks;
        const int tap = kb / p.cblocks;
        const int r = tap / p.S, s = tap - r * p.S;
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        const int cx = x0 * p.stride + s - p.pad_l, cy = y0 * p.stride + r - p.pad_t;
        uint8_t* b_hi = st + nops * A_TILE_BYTES;
        uint8_t* b_lo = b_hi + b_tile_bytes;
        const int nrow = n0 + (int)rank * b_rows;
        if (PAIR) {
          if (leader) mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * stage_bytes));
          tma_load_4d_2sm(&map_a_hi, &full_bar[stage], st, cb * p.kelems, cx, cy, img);
          if (p.split) tma_load_4d_2sm(&map_a_lo, &full_bar[stage], st + A_TILE_BYTES, cb * p.kelems, cx, cy, img);
          // weights: this CTA fetches slice (crank / 2) of its pair-half and multicasts it to the
          // same-parity CTA of every pair -- each weight byte crosses L2 -> SM once per cluster
          const int srows = b_rows / npairs;
          const int slice = (int)(crank >> 1);
          const int soff = slice * srows * 128;
          if (npairs == 1) {
            tma_load_2d_2sm(&map_b_hi, &full_bar[stage], b_hi, kb * p.kelems, nrow);
            if (p.split) tma_load_2d_2sm(&map_b_lo, &full_bar[stage], b_lo, kb * p.kelems, nrow);
          } else {
            tma_load_2d_2sm_mc(&map_b_hi, &full_bar[stage], b_hi + soff, kb * p.kelems, nrow + slice * srows, bcast_mask);
            if (p.split)
              tma_load_2d_2sm_mc(&map_b_lo, &full_bar[stage], b_lo + soff, kb * p.kelems, nrow + slice * srows, bcast_mask);
          }
        } else {
          mbar_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          tma_load_4d(&map_a_hi, &full_bar[stage], st, cb * p.kelems, cx, cy, img);
          tma_load_2d(&map_b_hi, &full_bar[stage], b_hi, kb * p.kelems, nrow);
          if (p.split) {
            tma_load_4d(&map_a_lo, &full_bar[stage], st + A_TILE_BYTES, cb * p.kelems, cx, cy, img);
            tma_load_2d(&map_b_lo, &full_bar[stage], b_lo, kb * p.kelems, nrow);
          }
        }
      }
    }
    __syncwarp();   // reconverge before the CTA / cluster barriers below
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only in pair mode) =====
    if (leader) {
      const uint32_t idesc = F16 ? make_idesc_f16(p.block_n, PAIR ? 256 : 128)
                                 : make_idesc_tf32(p.block_n, PAIR ? 256 : 128);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int stage = kb % nstages;
        const uint32_t parity = (kb / nstages) & 1;
        mbar_wait(&full_bar[stage], parity);
        tc_fence_after();
        if (kb == 0) TC_STAMP(2);
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t a_hi = st, a_lo = st + A_TILE_BYTES;
          const uint32_t b_hi = st + nops * A_TILE_BYTES, b_lo = b_hi + b_tile_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // 4 x (K = 8 tf32 = 32 B) per 128-byte swizzle row
            const uint32_t koff = k * 32;
            const uint32_t first = (kb | k) != 0;
            const uint64_t dah = make_sw128_desc(a_hi + koff), dal = make_sw128_desc(a_lo + koff);
            const uint64_t dbh = make_sw128_desc(b_hi + koff), dbl = make_sw128_desc(b_lo + koff);
            if (PAIR) {
              if (p.split) {
                if (F16) { umma_f16_2sm(tmem_base + acc2, dal, dbh, idesc, first); umma_f16_2sm(tmem_base + acc2, dah, dbl, idesc, 1); }
                else { umma_tf32_2sm(tmem_base + acc2, dal, dbh, idesc, first); umma_tf32_2sm(tmem_base + acc2, dah, dbl, idesc, 1); }
              }
              if (F16) umma_f16_2sm(tmem_base, dah, dbh, idesc, first);
              else umma_tf32_2sm(tmem_base, dah, dbh, idesc, first);
            } else {
              if (p.split) {
                if (F16) { umma_f16(tmem_base + acc2, dal, dbh, idesc, first); umma_f16(tmem_base + acc2, dah, dbl, idesc, 1); }
                else { umma_tf32(tmem_base + acc2, dal, dbh, idesc, first); umma_tf32(tmem_base + acc2, dah, dbl, idesc, 1); }
              }
              if (F16) umma_f16(tmem_base, dah, dbh, idesc, first);
              else umma_tf32(tmem_base, dah, dbh, idesc, first);
            }
          }
          // smem stage free (in both CTAs of a pair) once these MMAs retire
          if (PAIR) umma_commit_2sm(&empty_bar[stage], all_mask); else umma_commit(&empty_bar[stage]);
          if (kb == kblocks - 1) {
            if (PAIR) umma_commit_2sm(&tmem_full_bar, pair_mask); else umma_commit(&tmem_full_bar);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: warps 2..9.  A warp may only read the TMEM lane quarter (warp % 4); the two
    // warps of a quarter take alternate 32-column chunks.  The per-channel factors (operand scales
    // x folded BN scale, and the shift) are staged in shared memory WHILE the main loop runs: the
    // epilogue is not overlapped with MMAs (TMEM is full in the split modes), and fetching three
    // global values per output element inside it made it as long as the main loop itself
    // (measured: tile time = 61 us + 0.0625 us per MMA on the up3 layer before this change).
    const int quarter = warp & 3;
    const int egroup = (warp - 2) >> 2;              // 0 / 1: which chunks of 32 columns
    {
      const int e = (int)threadIdx.x - 64;           // 0..255 over the epilogue threads
      const int n = n0 + e;
      float mul = 0.f, add = 0.f;
      if (e < p.block_n && n < p.K) {
        mul = p.scale ? __ldg(p.scale + n) : 1.0f;
        if (F16) mul *= __ldg(p.act_inv) * __ldg(p.w_inv + n);      // powers of two: exact
        add = p.shift ? __ldg(p.shift + n) : 0.0f;
      }
      s_mul[e] = mul;
      s_add[e] = add;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    const int m = quarter * 32 + lane;               // accumulator row = box-order pixel index
    const int wx = m % p.wbox, hy = m / p.wbox;
    const int ox = x0 + wx, oy = y0 + hy;
    const bool pix_ok = (ox < p.Q) && (oy < p.P) && (img < p.N);
    const size_t pix = ((size_t)img * p.P + oy) * p.Q + ox;
    float amx = 0.0f;
    float s_out = 1.0f;
    if (F16 && p.out_hi) {
      const float xmax = p.in_amax ? __ldg(p.in_amax) : 32768.0f * __ldg(p.act_inv);
      const float bound = fmaf(xmax, p.bound_mul, p.bound_add);
      const unsigned bb = __float_as_uint(bound);
      int e = (int)((bb >> 23) & 0xffu) - 127;
      if (bb == 0u || !isfinite(bound)) e = 14;
      const int k = max(-100, min(100, 14 - e));
      s_out = __uint_as_float((unsigned)(127 + k) << 23);
      if (blockIdx.x == 0 && threadIdx.x == 64) {
        p.out_scal[0] = s_out;
        p.out_scal[1] = __uint_as_float((unsigned)(127 - k) << 23);
      }
    }
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    TC_STAMP(3);
    for (int c0 = egroup * 32; c0 < p.block_n; c0 += 64) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (p.split) {
        uint32_t v2[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc2 + (uint32_t)c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          v[j] = __float_as_uint(F16 ? fmaf(__uint_as_float(v2[j]), p.cross_scale, __uint_as_float(v[j]))
                                     : __uint_as_float(v[j]) + __uint_as_float(v2[j]));
      } else {
        tmem_ld_wait();
      }
      if (!p.out_nchw) {
        // transpose the 32 x 32 chunk through shared memory (the operand ring is idle now) so that
        // 8 lanes write one pixel's 32 channels = a full 128-byte line; storing straight from the
        // TMEM layout (one pixel row per lane) touched 32 lines with 16 bytes each per instruction
        float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 36);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * 36 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        __syncwarp();
        const int col = (lane & 7) * 4;
        const int n = n0 + c0 + col;
        const float4 mu = *reinterpret_cast<const float4*>(&s_mul[c0 + col]);
        const float4 ad = *reinterpret_cast<const float4*>(&s_add[c0 + col]);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int row = it * 4 + (lane >> 3);
          const unsigned long long rpix = __shfl_sync(0xffffffffu, (unsigned long long)pix, row);
          const int rok = __shfl_sync(0xffffffffu, (int)pix_ok, row);
          const float4 a = *reinterpret_cast<const float4*>(stg + row * 36 + col);
          if (rok && n < p.K && c0 + col < p.block_n) {
            float o[4] = {fmaf(a.x, mu.x, ad.x), fmaf(a.y, mu.y, ad.y), fmaf(a.z, mu.z, ad.z), fmaf(a.w, mu.w, ad.w)};
            float* dst = p.out + (size_t)rpix * p.K + n;
            if (n + 3 < p.K && (p.K & 3) == 0) {
              if (p.residual) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)rpix * p.K + n));
                o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
              }
              const float4 o4 = make_float4(tc_act(o[0], p.act), tc_act(o[1], p.act), tc_act(o[2], p.act), tc_act(o[3], p.act));
              amx = fmaxf(fmaxf(amx, fmaxf(fabsf(o4.x), fabsf(o4.y))), fmaxf(fabsf(o4.z), fabsf(o4.w)));
              if (p.out) *reinterpret_cast<float4*>(dst) = o4;
              if (F16 && p.out_hi) {
                // same arithmetic per element as f16_split_kernel
                uint2 h2, l2;
                split4_f16(o4.x * s_out, o4.y * s_out, o4.z * s_out, o4.w * s_out, h2, l2);
                const size_t o8 = ((size_t)rpix * p.K + n) >> 2;
                p.out_hi[o8] = h2;
                if (p.out_lo) p.out_lo[o8] = l2;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (n + q < p.K) {
                  float val = o[q];
                  if (p.residual) val += __ldg(p.residual + (size_t)rpix * p.K + n + q);
                  val = tc_act(val, p.act);
                  amx = fmaxf(amx, fabsf(val));
                  dst[q] = val;
                }
            }
          }
        }
        __syncwarp();
      } else if (pix_ok) {
        const size_t PQ = (size_t)p.P * p.Q;
        const size_t pp = (size_t)oy * p.Q + ox;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int nn = n0 + c0 + j;
          if (nn < p.K && c0 + j < p.block_n) {
            float val = fmaf(__uint_as_float(v[j]), s_mul[c0 + j], s_add[c0 + j]);
            if (p.residual) val += __ldg(p.residual + pix * p.K + nn);
            val = tc_act(val, p.act);
            amx = fmaxf(amx, fabsf(val));
            p.out[((size_t)img * p.K + nn) * PQ + pp] = val;
          }
        }
      }
    }
    if (p.amax_out) {
      amx = warp_max(amx);
      if (lane == 0 && amx > 0.0f) atomicMax(p.amax_out, __float_as_uint(amx));
    }
  }
  TC_STAMP(4);
  tc_fence_before();
  __syncthreads();
  TC_STAMP(5);
  if (PAIR) cluster_sync_all();       // no CTA exits while its peer's MMAs / TMA can still touch it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// x (* gate) -> hi = rna_tf32(x), lo = rna_tf32(x - hi); also used with lo == nullptr to pre-round
__global__ void __launch_bounds__(256) tf32_split_kernel(const float4* __restrict__ x,
                                                         const float* __restrict__ gate, int C,
                                                         long long hwc4, long long n4,
                                                         float4* __restrict__ hi,
                                                         float4* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(x + i);
    if (gate) {
      const int img = (int)(i / hwc4);
      const int c = (int)((i * 4) % C);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)img * C + c));
      v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    float4 h, l;
    uint32_t u;
#define CRESTE_SPLIT(f)                                                   \
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.f));                 \
    h.f = __uint_as_float(u);                                             \
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.f - h.f));           \
    l.f = __uint_as_float(u);
    CRESTE_SPLIT(x) CRESTE_SPLIT(y) CRESTE_SPLIT(z) CRESTE_SPLIT(w)
#undef CRESTE_SPLIT
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ---- 3xFP16 operand preparation.  amax over x (* gate) -> power-of-two scale s with
// amax * s in [2^14, 2^15); hi = fp16(x*s), lo = fp16((x*s - hi) * 2^11)  (both exact scalings).
// fp16 carries 11 significant bits like tf32, so hi*hi + (hi*lo + lo*hi) * 2^-11 has the accuracy of
// the 3xTF32 split at twice the tensor-pipe rate; the per-tensor scale keeps everything in fp16's
// normal range (small values go subnormal: absolute error <= 2^-25 * 2^-15 of the tensor maximum).
__global__ void __launch_bounds__(256) f16_amax_kernel(const float4* __restrict__ x, const float* __restrict__ gate,
                                                       int C, long long hwc4, long long n4,
                                                       unsigned* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(x + i);
    if (gate) {
      const int img = (int)(i / hwc4);
      const int c = (int)((i * 4) % C);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)img * C + c));
      v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}

// scal[0] = s (power of two), scal[1] = 1/s
__global__ void f16_scale_kernel(const unsigned* __restrict__ amax_bits, float* __restrict__ scal) {
  const unsigned b = *amax_bits;
  int e = (int)((b >> 23) & 0xffu) - 127;          // floor(log2(amax)) for normal amax
  if (b == 0u || !isfinite(__uint_as_float(b))) e = 14;
  const int k = max(-100, min(100, 14 - e));
  scal[0] = __uint_as_float((unsigned)(127 + k) << 23);
  scal[1] = __uint_as_float((unsigned)(127 - k) << 23);
}

// amax_bound != nullptr: the scale comes from an upper bound of max|x * gate| that travels with the tensor (the
// producing kernel's epilogue atomicMax, or the maximum over the inputs of an interpolation / concat): no amax pass
// and no scale kernel.  Any power-of-two scale with amax * s in [2^-4, 2^15] keeps the hi/lo split at 22 bits
// relative to the tensor maximum, so a loose bound costs nothing; block 0 publishes {s, 1/s} for the epilogue.
__global__ void __launch_bounds__(256) f16_split_kernel(const float4* __restrict__ x, const float* __restrict__ gate,
                                                        int C, long long hwc4, long long n4,
                                                        float* __restrict__ scal, const unsigned* __restrict__ amax_bound,
                                                        uint2* __restrict__ hi, uint2* __restrict__ lo) {
  float s;
  if (amax_bound) {
    const unsigned b = __ldg(amax_bound);
    int e = (int)((b >> 23) & 0xffu) - 127;
    if (b == 0u || !isfinite(__uint_as_float(b))) e = 14;
    const int k = max(-100, min(100, 14 - e));
    s = __uint_as_float((unsigned)(127 + k) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scal[0] = s; scal[1] = __uint_as_float((unsigned)(127 - k) << 23); }
  } else {
    s = __ldg(scal);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = __ldg(x + i);
    if (gate) {
      const int img = (int)(i / hwc4);
      const int c = (int)((i * 4) % C);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)img * C + c));
      v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
    }
    uint2 h2, l2;
    split4_f16(v.x * s, v.y * s, v.z * s, v.w * s, h2, l2);
    hi[i] = h2;
    if (lo) lo[i] = l2;
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)ptr;
  return fn;
}

// stride > 1: the box spans (wbox-1)*stride+1 input columns and TMA's element strides pick every
// stride-th pixel, so the tile still lands as wbox x hbox rows of 128 bytes
static int make_map_a(CUtensorMap* m, const void* base, int N, int H, int W, int C, int wbox, int hbox,
                      bool f16 = false, int stride = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CRESTE_ERR_NO_DEVICE; }
  const cuuint64_t eb = f16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
  cuuint32_t box[4] = {f16 ? 64u : 32u, (cuuint32_t)((wbox - 1) * stride + 1), (cuuint32_t)((hbox - 1) * stride + 1), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return CRESTE_ERR_ARG; }
  return 0;
}

static int make_map_b(CUtensorMap* m, const void* base, int ktot, int npad, int block_n, bool f16 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return CRESTE_ERR_NO_DEVICE; }
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)npad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * (f16 ? 2 : 4)};
  cuuint32_t box[2] = {f16 ? 64u : 32u, (cuuint32_t)block_n};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return CRESTE_ERR_ARG; }
  return 0;
}

// N tile: the whole (16-rounded) channel count when it fits one 256-column accumulator,
// otherwise the divisor-friendliest tile <= 256
static int pick_block_n(int K) {
  const int k16 = (K + 15) / 16 * 16;
  // experiment knob: cap the N tile (<= 128 lets two CTAs share an SM: 2 x 256 TMEM columns)
  int cap = 256;
  if (const char* e = getenv("CRESTE_TC_MAX_BN")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 16 == 0) cap = v; }
  if (k16 <= cap) return k16;
  const int tiles = (k16 + cap - 1) / cap;
  return ((k16 + tiles - 1) / tiles + 15) / 16 * 16;
}

static void pick_box(int P, int Q, int* wbox, int* hbox) {
  int best_w = 16;
  double best = -1.0;
  for (int w = 128; w >= 1; w >>= 1) {
    const int h = 128 / w;
    const double util = ((double)P * Q) / ((double)ceil_div(Q, w) * w * ceil_div(P, h) * h);
    // prefer squarer boxes on ties (fewer halo re-reads from L2)
    const double score = util - 1e-3 * fabs(log2((double)w / h) - 1.0);
    if (score > best) { best = score; best_w = w; }
  }
  *wbox = best_w;
  *hbox = 128 / best_w;
}

bool conv_tc_supported(const creste_conv_desc* d) {
  if (d->precision != 1 && d->precision != 2 && d->precision != 4 && d->precision != 5) return false;
  if ((d->stride != 1 && d->stride != 2) || d->C % 4 != 0 || d->K < 8) return false;
  if ((d->precision == 4 || d->precision == 5) && d->C % 8 != 0) return false;     // fp16 rows must be 16-byte multiples (TMA)
  if (d->R > 7 || d->S > 7) return false;
  if ((long long)d->N * d->P * d->Q < 128) return false;
  return true;
}

int conv_tc_layout(int K, int C, int R, int S, int* block_n, int* npad, int* cpad) {
  *block_n = pick_block_n(K);
  *npad = ceil_div(K, *block_n) * *block_n;
  *cpad = (C + 31) / 32 * 32;
  (void)R; (void)S;
  return 0;
}

size_t conv_tc_workspace_bytes(const creste_conv_desc* d) {
  const size_t n = (size_t)d->N * d->H * d->W * d->C * sizeof(float);
  if (d->precision == 4 || d->precision == 5) return 2 * align_up(n / 2, 1024) + 1024;     // fp16 hi, lo + scale scalars
  return d->precision == 1 ? 2 * align_up(n, 1024) : align_up(n, 1024);
}

static int conv_tc_main(const creste_conv_desc* d, float* x_hi, float* x_lo, float* scal, const float* w_packed,
                        const float* scale, const float* shift, const float* residual, float* out, unsigned* amax_out,
                        const TcSplitOut* so, cudaStream_t st);

int conv_tc_launch(const creste_conv_desc* d, const float* x, const float* w_packed, const float* scale,
                   const float* shift, const float* gate, const float* residual, float* out, const float* amax_in,
                   unsigned* amax_out, void* ws, size_t ws_bytes, cudaStream_t st, const TcSplitOut* so) {
  if (ws_bytes < conv_tc_workspace_bytes(d) || !ws) {
    set_error("creste_conv2d(tc): workspace %zu < %zu", ws_bytes, conv_tc_workspace_bytes(d));
    return CRESTE_ERR_WORKSPACE;
  }
  // precision 4 = 3xFP16 (hi/lo split, fp32-faithful); 5 = single-pass fp16 (hi only: the arithmetic class of the
  // reference's own default GPU run, cuDNN TF32 -- 11 significant bits per operand; reported, never asserted)
  const bool f16 = d->precision == 4 || d->precision == 5;
  const int split = d->precision == 1 || d->precision == 4;
  int block_n, npad, cpad;
  conv_tc_layout(d->K, d->C, d->R, d->S, &block_n, &npad, &cpad);
  if (f16) cpad = (d->C + 63) / 64 * 64;
  const int ktot = d->R * d->S * cpad;
  const size_t numel = (size_t)d->N * d->H * d->W * d->C;
  float* x_hi = (float*)ws;
  float* x_lo = split ? (float*)((char*)ws + align_up(numel * (f16 ? 2 : 4), 1024)) : nullptr;
  float* scal = nullptr;
  if (f16) {
    scal = (float*)((char*)ws + 2 * align_up(numel * 2, 1024));       // [s, 1/s, amax bits, -]
    unsigned* amax = (unsigned*)(scal + 2);
    const long long n4 = (long long)(numel / 4);
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    const long long hwc4 = (long long)d->H * d->W * d->C / 4;
    int rc;
    if (!amax_in) {          // no bound travels with the tensor: one extra pass over it
      CRESTE_CUDA(cudaMemsetAsync(amax, 0, 4, st));
      f16_amax_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)x, gate, d->C, hwc4, n4, amax);
      rc = launch_check("f16_amax_kernel");
      if (rc) return rc;
      f16_scale_kernel<<<1, 1, 0, st>>>(amax, scal);
      rc = launch_check("f16_scale_kernel");
      if (rc) return rc;
    }
    f16_split_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)x, gate, d->C, hwc4, n4, scal,
                                                  (const unsigned*)amax_in, (uint2*)x_hi, (uint2*)x_lo);
    rc = launch_check("f16_split_kernel");
    if (rc) return rc;
  } else {
    const long long n4 = (long long)(numel / 4);
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    tf32_split_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)x, gate, d->C,
                                                   (long long)d->H * d->W * d->C / 4, n4, (float4*)x_hi,
                                                   (float4*)x_lo);
    int rc = launch_check("tf32_split_kernel");
    if (rc) return rc;
  }
  return conv_tc_main(d, x_hi, x_lo, scal, w_packed, scale, shift, residual, out, amax_out, so, st);
}

// operands already split (by the pre-pass above, or written that way by the producing kernel)
int conv_tc_presplit_launch(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                            const float* w_packed, const float* scale, const float* shift, const float* residual,
                            float* out, unsigned* amax_out, cudaStream_t st, const TcSplitOut* so) {
  return conv_tc_main(d, (float*)x_hi, (float*)x_lo, (float*)x_scal, w_packed, scale, shift, residual, out, amax_out, so, st);
}

static int conv_tc_main(const creste_conv_desc* d, float* x_hi, float* x_lo, float* scal, const float* w_packed,
                        const float* scale, const float* shift, const float* residual, float* out, unsigned* amax_out,
                        const TcSplitOut* so, cudaStream_t st) {
  const bool f16 = d->precision == 4 || d->precision == 5;
  if (so && so->hi) {
    if (!f16 || d->out_nchw || d->K % 8 != 0 || !so->scal || ((uintptr_t)so->hi & 7u) || ((uintptr_t)so->lo & 7u)) {
      set_error("creste_conv2d(tc): the split output needs a 3xFP16 / fp16 mode, NHWC output and K %% 8 == 0");
      return CRESTE_ERR_ARG;
    }
  } else if (!out) {
    set_error("creste_conv2d(tc): no output");
    return CRESTE_ERR_ARG;
  }
  const int split = d->precision == 1 || d->precision == 4;
  int block_n, npad, cpad;
  conv_tc_layout(d->K, d->C, d->R, d->S, &block_n, &npad, &cpad);
  if (f16) cpad = (d->C + 63) / 64 * 64;
  const int ktot = d->R * d->S * cpad;
  TcParams p;
  p.scale = scale; p.shift = shift; p.residual = residual; p.out = out;
  p.N = d->N; p.P = d->P; p.Q = d->Q; p.K = d->K;
  p.R = d->R; p.S = d->S; p.pad_t = d->pad_t; p.pad_l = d->pad_l; p.stride = d->stride;
  p.cblocks = cpad / 32;
  pick_box(d->P, d->Q, &p.wbox, &p.hbox);
  p.tiles_x = ceil_div(d->Q, p.wbox);
  p.tiles_y = ceil_div(d->P, p.hbox);
  p.block_n = block_n; p.act = d->act; p.split = split; p.out_nchw = d->out_nchw;
  p.amax_out = amax_out;
  p.out_hi = so ? (uint2*)so->hi : nullptr;
  p.out_lo = so ? (uint2*)so->lo : nullptr;
  p.out_scal = so ? so->scal : nullptr;
  p.bound_mul = so ? so->bound_mul : 0.0f;
  p.bound_add = so ? so->bound_add : 0.0f;
  p.in_amax = so ? so->in_amax : nullptr;

  // two CTAs on adjacent M tiles form a tcgen05 CTA pair (cta_group::2, M = 256)
  const int m_tiles = d->N * p.tiles_y * p.tiles_x;
  // Low-resolution layers (the deep MBConv stages, the BEV decoder's layer 3 / 4; every layer at B = 1) have a
  // handful of M tiles: 4 - 32 CTAs each walking the whole K loop while 120 SMs idle.  Their N tile is cut into
  // f equal slices (a divisor of the packed tile, so the weight layout does not change; >= 32 channels, multiples of
  // 32 so that a CTA pair still splits whole swizzle atoms): f x as many CTAs, each k-step f x shorter.
  if (!getenv("CRESTE_TC_NO_SHRINK")) {
    const int ctas = m_tiles * (npad / block_n);
    if (ctas < 96) {
      int best = 1;
      for (int f = 2; f <= 8; ++f) {
        if (block_n % f || (block_n / f) % 32 || block_n / f < 32) continue;
        best = f;
        if (ctas * f >= 120) break;
      }
      block_n /= best;
      p.block_n = block_n;
    }
  }
  // cluster = 1 (no pairing), 2 (one CTA pair, the default) or 4 / 8 (CRESTE_TC_CLUSTER: 2 / 4 pairs
  // sharing the weight tile by TMA multicast; each slice keeps whole 8-row swizzle atoms).
  // Measured on the up3 conv: multicast cuts the L2->SM weight traffic by 25 % / 37 % but the
  // per-SM rate does not move, and 4- / 8-CTA clusters only fill 132 / 120 of the 148 SMs
  // (3.24 -> 3.62 / 3.95 ms) -- the kernel is not L2-bandwidth bound, so pairs stay the default.
  int cl = (m_tiles >= 2 && (block_n % 16) == 0 && !getenv("CRESTE_TC_NO_PAIR")) ? 2 : 1;
  if (cl == 2) {
    int want = 2;
    if (const char* e = getenv("CRESTE_TC_CLUSTER")) want = atoi(e);
    while (want > 2 && !(m_tiles >= 2 * want && (block_n / want) % 8 == 0)) want >>= 1;
    if (want == 4 || want == 8) cl = want;
  }
  p.cl = cl;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map_a(&ma_hi, x_hi, d->N, d->H, d->W, d->C, p.wbox, p.hbox, f16, d->stride))) return rc;
  if ((rc = make_map_a(&ma_lo, split ? x_lo : x_hi, d->N, d->H, d->W, d->C, p.wbox, p.hbox, f16, d->stride))) return rc;
  // packed weights: [hi][lo] (fp32 words for tf32; fp16 halves for 3xFP16, followed by w_inv[npad] fp32)
  const size_t wel = (size_t)npad * ktot;
  const void* w_hi = w_packed;
  const void* w_lo = f16 ? (const void*)((const char*)w_packed + wel * 2) : (const void*)(w_packed + wel);
  if ((rc = make_map_b(&mb_hi, w_hi, ktot, npad, block_n / cl, f16))) return rc;
  if ((rc = make_map_b(&mb_lo, split ? w_lo : w_hi, ktot, npad, block_n / cl, f16))) return rc;
  p.kelems = f16 ? 64 : 32;
  p.cblocks = cpad / p.kelems;
  p.act_inv = f16 ? scal + 1 : nullptr;
  p.w_inv = f16 ? (const float*)((const char*)w_packed + wel * 4) : nullptr;
  p.cross_scale = f16 ? (1.0f / 2048.0f) : 1.0f;

  const int nops = split ? 2 : 1;
  const size_t stage_bytes = (size_t)nops * (A_TILE_BYTES + (block_n / (cl >= 2 ? 2 : 1)) * 128);
  // N tiles of <= 128 channels need <= 256 TMEM columns (main + cross accumulators): TWO CTAs fit on an SM if each
  // keeps its operand ring under ~100 KB, and then one CTA's prologue / epilogue (TMEM -> registers -> transpose ->
  // global, now including the fp16 split) runs under the other's MMAs.  One tile per CTA with the whole SM to itself
  // left the tensor pipe idle for ~38 % of the tile time on these layers (profiles/r1b_conv_tc_bottleneck.md).
  // CRESTE_TC_SMEM_KB overrides the per-CTA ring budget (200 = one CTA per SM, the round-1 behaviour).
  int ring_kb = block_n <= 128 ? 100 : 200;
  if (const char* e = getenv("CRESTE_TC_SMEM_KB")) { const int v = atoi(e); if (v >= 64 && v <= 200) ring_kb = v; }
  int nstages = (int)(((size_t)ring_kb * 1024) / stage_bytes);
  if (nstages > 8) nstages = 8;
  if (nstages < 2) { set_error("creste_conv2d(tc): stage too large"); return CRESTE_ERR_ARG; }
  const size_t smem = (size_t)nstages * stage_bytes + 1024;
  auto kern = f16 ? (cl >= 2 ? conv_tc_kernel<true, true> : conv_tc_kernel<false, true>)
                  : (cl >= 2 ? conv_tc_kernel<true, false> : conv_tc_kernel<false, false>);
  CRESTE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.n_tiles = npad / block_n;
  dim3 grid(ceil_div(m_tiles, cl) * cl * p.n_tiles, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CRESTE_CUDA(cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, p, nstages));
  return launch_check("conv_tc_kernel");
}


// =====================================================================================================
// Weight gradient of the stride-1 dense convolutions on the tensor cores (stage-1 / stage-3 training).
//
//   dw[tap][c][k] = sum over output pixels of  g[pix][k] * x[pix shifted by the tap][c]
//
// GEMM view per filter tap: D[M = 128 output channels][N = BN input channels] accumulated over the
// PIXELS.  Both operands are therefore "MN-major": the reduction index (pixel) is the slow axis of the
// NHWC tensors.  A TMA box {64 channels, wbox, hbox, 1} lands as 64 pixel rows x 128 bytes with the
// 128-byte swizzle -- the canonical MN-major SWIZZLE_128B UMMA layout ((T,8,m),(8,k)) with
// SBO = 1024 B between 8-pixel groups and LBO = one box (8192 B) between 64-channel groups; the
// instruction descriptor carries the a_major / b_major transpose bits.  The shifted x box (tap offset,
// zero padding, ragged edges, channel tails) is TMA out-of-bounds zero fill, exactly as in the forward.
// Precision: 3xFP16 split of both operands (same pre-pass as the forward), main and cross terms in
// separate TMEM accumulators.  One CTA = (128-k tile, BN-c tile, tap, pixel split); partial tiles
// part[split][tap][C][K] are summed in a fixed order by wg_reduce_kernel.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (one TMEM lane quarter each).
struct WgParams {
  float* part;                 // [splits][R*S][C][K]
  const float* sx; const float* sg;     // operand scale records {s, 1/s}
  int N, P, Q, C, K, R, S, pad_t, pad_l;
  int wbox, hbox, tiles_x, tiles_y, ntiles;
  int bn, nb;                  // input-channel tile (multiple of 64) and its 64-channel box count
  int m_tiles, n_tiles, splits;
  int tg, ngroups;             // filter taps per CTA (they share the g tile) and tap groups = ceil(R*S / tg)
  uint32_t lbo, sbo;           // descriptor offsets in bytes
};
constexpr int WG_THREADS = 192;
constexpr int WG_BOX_BYTES = 64 * 128;       // 64 pixels x 64 fp16 channels

__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                WgParams p, int nstages) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int split = t % p.splits; t /= p.splits;
  const int group = t % p.ngroups; t /= p.ngroups;
  const int n_tile = t % p.n_tiles;
  const int m_tile = t / p.n_tiles;
  // Narrow layers (input-channel tile <= 128: accumulators of <= 2 x 128 TMEM columns per tap) put several filter
  // taps in one CTA: the g tile of a pixel block is fetched once for all of them and only the shifted x boxes differ.
  // One CTA per tap moved 25 x (g + x) through L2 for the reward FCN's 5x5 layer: L2-bandwidth bound at 89 TFLOP/s.
  const int tap0 = group * p.tg;
  const int gcount = min(p.tg, p.R * p.S - tap0);
  const int m0 = m_tile * 128, n0 = n_tile * p.bn;
  const int per = (p.ntiles + p.splits - 1) / p.splits;
  const int t0 = split * per, t1 = min(t0 + per, p.ntiles);
  const int iters = max(t1 - t0, 0);
  const int a_bytes = 2 * WG_BOX_BYTES, b_bytes = p.nb * WG_BOX_BYTES;
  const int stage_bytes = 2 * (a_bytes + p.tg * b_bytes);
  // 64-channel boxes that lie entirely beyond K / C are not fetched: their accumulator rows / columns hold
  // garbage that the epilogue never stores (rows and columns of D are independent)
  const int a_boxes = (p.K - m0 > 64) ? 2 : 1;
  const int b_boxes = min(p.nb, (p.C - n0 + 63) / 64);
  const uint32_t tx_bytes = (uint32_t)(2 * (a_boxes + gcount * b_boxes) * WG_BOX_BYTES);
  const uint32_t acc_cols = p.bn <= 64 ? 64 : (p.bn <= 128 ? 128 : 256);
  const uint32_t tmem_cols = p.tg * 2 * acc_cols;      // 128 / 256 / 512: a power of two by construction of tg

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_g_hi); prefetch_tmap(&map_g_lo); prefetch_tmap(&map_x_hi); prefetch_tmap(&map_x_lo);
    for (int i = 0; i < nstages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int stage = it % nstages;
        mbar_wait(&empty_bar[stage], ((it / nstages) & 1) ^ 1);
        int tt = t0 + it;
        const int tx = tt % p.tiles_x; tt /= p.tiles_x;
        const int ty = tt % p.tiles_y;
        const int img = tt / p.tiles_y;
        const int q0 = tx * p.wbox, p0 = ty * p.hbox;
        uint8_t* st = smem + (size_t)stage * stage_bytes;
        uint8_t* a_hi = st; uint8_t* a_lo = st + a_bytes;
        mbar_expect_tx(&full_bar[stage], tx_bytes);
        for (int j = 0; j < a_boxes; ++j) {
          tma_load_4d(&map_g_hi, &full_bar[stage], a_hi + j * WG_BOX_BYTES, m0 + j * 64, q0, p0, img);
          tma_load_4d(&map_g_lo, &full_bar[stage], a_lo + j * WG_BOX_BYTES, m0 + j * 64, q0, p0, img);
        }
        for (int tj = 0; tj < gcount; ++tj) {
          const int tap = tap0 + tj;
          const int r = tap / p.S, s = tap - r * p.S;
          uint8_t* b_hi = st + 2 * a_bytes + (size_t)tj * 2 * b_bytes; uint8_t* b_lo = b_hi + b_bytes;
          const int cx = q0 + s - p.pad_l, cy = p0 + r - p.pad_t;
          for (int j = 0; j < b_boxes; ++j) {
            tma_load_4d(&map_x_hi, &full_bar[stage], b_hi + j * WG_BOX_BYTES, n0 + j * 64, cx, cy, img);
            tma_load_4d(&map_x_lo, &full_bar[stage], b_lo + j * WG_BOX_BYTES, n0 + j * 64, cx, cy, img);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // fp16 A/B, fp32 accumulate, M = 128, N = bn, A and B MN-major (transpose bits 15 / 16)
    const uint32_t idesc = make_idesc_f16(p.bn, 128) | (1u << 15) | (1u << 16);
    for (int it = 0; it < iters; ++it) {
      const int stage = it % nstages;
      mbar_wait(&full_bar[stage], (it / nstages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + a_bytes;
        for (int tj = 0; tj < gcount; ++tj) {
          const uint32_t b_hi = st + 2 * a_bytes + (uint32_t)tj * 2 * b_bytes, b_lo = b_hi + b_bytes;
          const uint32_t acc = tmem_base + (uint32_t)tj * 2 * acc_cols;
#pragma unroll
          for (int k = 0; k < 4; ++k) {          // 4 x 16 pixels per 64-pixel box
            const uint32_t koff = k * 16 * 128;
            const uint32_t first = (it | k) != 0;
            const uint64_t dah = make_mn_sw128_desc(a_hi + koff, p.lbo, p.sbo), dal = make_mn_sw128_desc(a_lo + koff, p.lbo, p.sbo);
            const uint64_t dbh = make_mn_sw128_desc(b_hi + koff, p.lbo, p.sbo), dbl = make_mn_sw128_desc(b_lo + koff, p.lbo, p.sbo);
            umma_f16(acc + acc_cols, dal, dbh, idesc, first);
            umma_f16(acc + acc_cols, dah, dbl, idesc, 1);
            umma_f16(acc, dah, dbh, idesc, first);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (it == iters - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps 2..5: TMEM lane quarter = warp % 4; lane = output channel k, columns = input channel c
    const int quarter = warp & 3;
    const int k = m0 + quarter * 32 + lane;
    if (iters > 0) {
      mbar_wait(&tmem_full_bar, 0);
      tc_fence_after();
    }
    const float inv = __ldg(p.sx + 1) * __ldg(p.sg + 1);
    const float cross = 1.0f / 2048.0f;
    for (int tj = 0; tj < gcount; ++tj) {
      float* dst = p.part + ((size_t)split * (p.R * p.S) + tap0 + tj) * (size_t)p.C * p.K;
      for (int c0 = 0; c0 < p.bn; c0 += 32) {
        uint32_t vm[32], vc[32];
        if (iters > 0) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)tj * 2 * acc_cols + (uint32_t)c0;
          tmem_ld32(taddr, vm);
          tmem_ld32(taddr + acc_cols, vc);
          tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = n0 + c0 + j;
          if (c < p.C && k < p.K) {
            const float v = iters > 0 ? (__uint_as_float(vm[j]) + __uint_as_float(vc[j]) * cross) * inv : 0.0f;
            dst[(size_t)c * p.K + k] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__global__ void __launch_bounds__(256) wg_reduce_kernel(const float* __restrict__ part, int rows, long long n,
                                                        float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float tot = 0.f;
  for (int j = 0; j < rows; ++j) tot += __ldg(part + (size_t)j * n + i);
  out[i] = tot;
}

// fp32 NHWC tensor -> fp16 hi / lo halves + {s, 1/s} (the forward's 3xFP16 operand pre-pass)
static int wg_split(const float* x, size_t numel, int C, void* hi, void* lo, float* scal, cudaStream_t st) {
  unsigned* amax = (unsigned*)(scal + 2);
  CRESTE_CUDA(cudaMemsetAsync(amax, 0, 4, st));
  const long long n4 = (long long)(numel / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f16_amax_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)x, nullptr, C, 1, n4, amax);
  int rc = launch_check("f16_amax_kernel");
  if (rc) return rc;
  f16_scale_kernel<<<1, 1, 0, st>>>(amax, scal);
  if ((rc = launch_check("f16_scale_kernel"))) return rc;
  f16_split_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)x, nullptr, C, 1, n4, scal, nullptr, (uint2*)hi, (uint2*)lo);
  return launch_check("f16_split_kernel");
}

// the same with max|x| already known (published by the kernel that produced x): no amax pass, no scale kernel
static int wg_split_known(const float* x, size_t numel, int C, void* hi, void* lo, float* scal, const float* amax,
                          cudaStream_t st) {
  const long long n4 = (long long)(numel / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f16_split_kernel<<<(int)blocks, 256, 0, st>>>((const float4*)x, nullptr, C, 1, n4, scal, (const unsigned*)amax,
                                                (uint2*)hi, (uint2*)lo);
  return launch_check("f16_split_kernel");
}

static void wg_pick_box(int P, int Q, int* wbox, int* hbox) {
  int best_w = 8;
  double best = -1.0;
  for (int w = 64; w >= 1; w >>= 1) {
    const int h = 64 / w;
    const double util = ((double)P * Q) / ((double)ceil_div(Q, w) * w * ceil_div(P, h) * h);
    const double score = util - 1e-3 * fabs(log2((double)w / h));
    if (score > best) { best = score; best_w = w; }
  }
  *wbox = best_w;
  *hbox = 64 / best_w;
}

struct WgPlan { int bn, nb, m_tiles, n_tiles, wbox, hbox, tiles_x, tiles_y, ntiles, splits, tg, ngroups; size_t nx, ng; };

static WgPlan wg_plan(const creste_conv_desc* d) {
  WgPlan w;
  const int c64 = (d->C + 63) / 64 * 64;
  const int nt = (c64 + 255) / 256;
  w.bn = ((c64 / 64 + nt - 1) / nt) * 64;          // <= 256, balanced over the N tiles
  w.nb = w.bn / 64;
  w.n_tiles = ceil_div(d->C, w.bn);
  w.m_tiles = ceil_div(d->K, 128);
  wg_pick_box(d->P, d->Q, &w.wbox, &w.hbox);
  w.tiles_x = ceil_div(d->Q, w.wbox);
  w.tiles_y = ceil_div(d->P, w.hbox);
  w.ntiles = d->N * w.tiles_y * w.tiles_x;
  // pixel splits: minimise (waves over the 148 SMs) x (pixel tiles per CTA + a fixed per-CTA cost of ~6 tiles for
  // the prologue / epilogue), so that the grid does not spill a nearly empty last wave (25 taps x 12 splits = 300
  // CTAs ran as three waves on the reward FCN's 5x5 layers)
  // taps per CTA: as many as the 512 TMEM columns hold main + cross accumulators for (4 / 2 / 1 for input-channel
  // tiles of 64 / 128 / 256), rounded down to a power of two so that the TMEM allocation is one
  int tg = w.bn <= 64 ? 4 : (w.bn <= 128 ? 2 : 1);
  if (const char* e = getenv("CRESTE_WGRAD_TAPS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) tg = v < tg ? v : tg; }
  while (tg > 1 && tg > d->R * d->S) tg >>= 1;
  w.tg = tg;
  w.ngroups = ceil_div(d->R * d->S, tg);
  const int base = w.m_tiles * w.n_tiles * w.ngroups;
  const int max_splits = w.ntiles / 8 > 1 ? w.ntiles / 8 : 1;
  int best_s = 1;
  long long best_cost = -1;
  for (int sp = 1; sp <= max_splits && (long long)base * sp <= 4 * 148 + base; ++sp) {
    const long long waves = ceil_div(base * sp, 148);
    const long long cost = waves * (ceil_div(w.ntiles, sp) + 6);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = sp; }
  }
  w.splits = best_s;
  w.nx = (size_t)d->N * d->H * d->W * d->C;
  w.ng = (size_t)d->N * d->P * d->Q * d->K;
  return w;
}

bool wgrad_tc_supported(const creste_conv_desc* d) {
  if (d->stride != 1 || d->C % 8 != 0 || d->K % 8 != 0 || d->C < 8 || d->K < 8) return false;
  if (d->R > 7 || d->S > 7) return false;
  if ((long long)d->N * d->P * d->Q < 512) return false;
  return true;
}

size_t wgrad_tc_workspace_bytes(const creste_conv_desc* d) {
  const WgPlan w = wg_plan(d);
  return 2 * align_up(w.nx * 2, 1024) + 2 * align_up(w.ng * 2, 1024) + 1024 +
         (size_t)w.splits * d->R * d->S * d->C * d->K * sizeof(float);
}

int f16_split_known_launch(const float* x, size_t numel, void* hi, void* lo, float* scal, const float* amax,
                           cudaStream_t st) {
  return wg_split_known(x, numel, 4, hi, lo, scal, amax, st);
}

static int wgrad_tc_core(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                         const void* g_hi, const void* g_lo, const float* g_scal, float* part, float* dw, cudaStream_t st);

int wgrad_tc_launch(const creste_conv_desc* d, const float* x, const float* g, float* dw, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  if (!wgrad_tc_supported(d)) { set_error("creste_conv2d_wgrad_tc: shape not served (C, K multiples of 8, stride 1, >= 512 output pixels)"); return CRESTE_ERR_ARG; }
  if (!ws || ws_bytes < wgrad_tc_workspace_bytes(d)) { set_error("creste_conv2d_wgrad_tc: workspace"); return CRESTE_ERR_WORKSPACE; }
  const WgPlan w = wg_plan(d);
  char* base = (char*)ws;
  void* x_hi = base; void* x_lo = base + align_up(w.nx * 2, 1024);
  char* gb = base + 2 * align_up(w.nx * 2, 1024);
  void* g_hi = gb; void* g_lo = gb + align_up(w.ng * 2, 1024);
  float* scal = (float*)(gb + 2 * align_up(w.ng * 2, 1024));      // x: [0..3], g: [4..7]
  float* part = (float*)((char*)scal + 1024);
  int rc;
  if ((rc = wg_split(x, w.nx, d->C, x_hi, x_lo, scal, st))) return rc;
  if ((rc = wg_split(g, w.ng, d->K, g_hi, g_lo, scal + 4, st))) return rc;
  return wgrad_tc_core(d, x_hi, x_lo, scal, g_hi, g_lo, scal + 4, part, dw, st);
}

// operands already split (the forward conv's saved operand, the output gradient split once for the data and the
// weight gradient): only the partial-tile region of the workspace is used
int wgrad_tc_presplit_launch(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                             const void* g_hi, const void* g_lo, const float* g_scal, float* dw, void* ws,
                             size_t ws_bytes, cudaStream_t st) {
  if (!wgrad_tc_supported(d)) { set_error("creste_conv2d_wgrad_tc_presplit: shape not served"); return CRESTE_ERR_ARG; }
  if (!ws || ws_bytes < wgrad_tc_workspace_bytes(d)) { set_error("creste_conv2d_wgrad_tc_presplit: workspace"); return CRESTE_ERR_WORKSPACE; }
  const WgPlan w = wg_plan(d);
  float* part = (float*)((char*)ws + 2 * align_up(w.nx * 2, 1024) + 2 * align_up(w.ng * 2, 1024) + 1024);
  return wgrad_tc_core(d, x_hi, x_lo, x_scal, g_hi, g_lo, g_scal, part, dw, st);
}

// amax -> power-of-two scale -> fp16 hi / lo of a dense fp32 tensor (the pre-pass of creste_conv2d, stand-alone);
// scal = DEVICE float[4]: {s, 1/s, amax bits, -}
int f16_split_launch(const float* x, size_t numel, void* hi, void* lo, float* scal, cudaStream_t st) {
  return wg_split(x, numel, 4, hi, lo, scal, st);
}

static int wgrad_tc_core(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* scal_x,
                         const void* g_hi, const void* g_lo, const float* scal_g, float* part, float* dw, cudaStream_t st) {
  const WgPlan w = wg_plan(d);
  int rc;
  CUtensorMap mg_hi, mg_lo, mx_hi, mx_lo;
  // box {64 channels, wbox, hbox, 1}: make_map_a's f16 form with a 64-pixel box
  if ((rc = make_map_a(&mg_hi, g_hi, d->N, d->P, d->Q, d->K, w.wbox, w.hbox, true, 1))) return rc;
  if ((rc = make_map_a(&mg_lo, g_lo, d->N, d->P, d->Q, d->K, w.wbox, w.hbox, true, 1))) return rc;
  if ((rc = make_map_a(&mx_hi, x_hi, d->N, d->H, d->W, d->C, w.wbox, w.hbox, true, 1))) return rc;
  if ((rc = make_map_a(&mx_lo, x_lo, d->N, d->H, d->W, d->C, w.wbox, w.hbox, true, 1))) return rc;
  WgParams p;
  p.part = part; p.sx = scal_x; p.sg = scal_g;
  p.N = d->N; p.P = d->P; p.Q = d->Q; p.C = d->C; p.K = d->K; p.R = d->R; p.S = d->S;
  p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  p.wbox = w.wbox; p.hbox = w.hbox; p.tiles_x = w.tiles_x; p.tiles_y = w.tiles_y; p.ntiles = w.ntiles;
  p.bn = w.bn; p.nb = w.nb; p.m_tiles = w.m_tiles; p.n_tiles = w.n_tiles; p.splits = w.splits;
  p.tg = w.tg; p.ngroups = w.ngroups;
  p.lbo = WG_BOX_BYTES; p.sbo = 1024;
  if (const char* e = getenv("CRESTE_WGRAD_SWAP")) if (atoi(e)) { p.lbo = 1024; p.sbo = WG_BOX_BYTES; }
  const size_t stage_bytes = (size_t)2 * (2 + w.tg * w.nb) * WG_BOX_BYTES;
  int nstages = (int)((200 * 1024) / stage_bytes);
  if (nstages > 8) nstages = 8;
  if (nstages < 2) { set_error("creste_conv2d_wgrad_tc: stage too large"); return CRESTE_ERR_ARG; }
  const size_t smem = (size_t)nstages * stage_bytes + 1024;
  CRESTE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = w.m_tiles * w.n_tiles * w.ngroups * w.splits;
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(mg_hi, mg_lo, mx_hi, mx_lo, p, nstages);
  if ((rc = launch_check("wgrad_tc_kernel"))) return rc;
  const long long n = (long long)d->R * d->S * d->C * d->K;
  wg_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(part, w.splits, n, dw);
  return launch_check("wg_reduce_kernel");
}


// ---- 3xFP16 weight operand in ONE launch (training: the weights change every step, so the ~20 small
// torch kernels of the cached inference pack would be paid per conv per step).  Logical weights
// w[k][c][r][s] are read through element strides (the data-gradient conv passes the transposed view of the
// flipped filter); block = one output channel: amax -> power-of-two scale (amax * s in [2^14, 2^15)) ->
// hi = fp16(w*s), lo = fp16((w*s - hi) * 2^11) in the [npad][R*S][cpad] layout, inv[k] = 1/s.
__global__ void __launch_bounds__(256) pack_weight_f16_kernel(const float* __restrict__ w, long long sK, long long sC,
                                                              long long sR, long long sS, int K, int C, int R, int S,
                                                              int cpad, __half* __restrict__ hi, __half* __restrict__ lo,
                                                              float* __restrict__ inv) {
  __shared__ float s_red[8];
  __shared__ float s_scale;
  const int k = blockIdx.x;
  const int RS = R * S;
  const int n = RS * C;
  float m = 0.f;
  if (k < K)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int c = i % C, t = i / C;
      const int r = t / S, s = t - r * S;
      m = fmaxf(m, fabsf(__ldg(w + k * sK + c * sC + r * sR + s * sS)));
    }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = s_red[0];
    for (int i = 1; i < 8; ++i) a = fmaxf(a, s_red[i]);
    const unsigned b = __float_as_uint(a);
    int e = (int)((b >> 23) & 0xffu) - 127;
    if (b == 0u || !isfinite(a)) e = 0;
    const int kk = max(-100, min(100, 14 - e));
    s_scale = __uint_as_float((unsigned)(127 + kk) << 23);
    inv[k] = __uint_as_float((unsigned)(127 - kk) << 23);
  }
  __syncthreads();
  const float sc = s_scale;
  const size_t row = (size_t)k * RS * cpad;
  for (int i = threadIdx.x; i < RS * cpad; i += blockDim.x) {
    const int c = i % cpad, t = i / cpad;
    float v = 0.f;
    if (k < K && c < C) {
      const int r = t / S, s = t - r * S;
      v = __ldg(w + k * sK + c * sC + r * sR + s * sS) * sc;
    }
    const __half h = __float2half_rn(v);
    hi[row + i] = h;
    lo[row + i] = __float2half_rn((v - __half2float(h)) * 2048.0f);
  }
}

int pack_weight_f16_launch(const float* w, long long sK, long long sC, long long sR, long long sS, int K, int C, int R,
                           int S, float* out, cudaStream_t st) {
  int block_n, npad, cpad32;
  conv_tc_layout(K, C, R, S, &block_n, &npad, &cpad32);
  const int cpad = (C + 63) / 64 * 64;
  const size_t wel = (size_t)npad * R * S * cpad;
  __half* hi = (__half*)out;
  __half* lo = hi + wel;
  float* inv = (float*)((char*)out + wel * 4);
  pack_weight_f16_kernel<<<npad, 256, 0, st>>>(w, sK, sC, sR, sS, K, C, R, S, cpad, hi, lo, inv);
  return launch_check("pack_weight_f16_kernel");
}

}  // namespace creste

/* development aid: per-CTA globaltimer stamps of the next tensor-core conv launches (8 u64 per CTA) */
extern "C" int creste_conv2d_tc_debug(void* dev_buf) {
  unsigned long long* p = (unsigned long long*)dev_buf;
  return (int)cudaMemcpyToSymbol(creste::g_tc_dbg, &p, sizeof(p));
}

extern "C" int creste_conv2d_tc_supported(const creste_conv_desc* d) {
  return d && creste::conv_tc_supported(d) ? 1 : 0;
}

// weight layout helper for the host side (Python packs on the GPU with torch; this reports sizes)
extern "C" int creste_conv2d_tc_layout(int K, int C, int R, int S, int* block_n, int* npad, int* cpad) {
  return creste::conv_tc_layout(K, C, R, S, block_n, npad, cpad);
}

/* tcgen05 weight gradient (3xFP16): dw [R*S*C][K] like creste_conv2d_wgrad, for C, K >= 64. */
extern "C" int creste_conv2d_wgrad_tc_supported(const creste_conv_desc* d) {
  return d && creste::wgrad_tc_supported(d) ? 1 : 0;
}
extern "C" size_t creste_conv2d_wgrad_tc_workspace_bytes(const creste_conv_desc* d) {
  return d ? creste::wgrad_tc_workspace_bytes(d) : 0;
}
extern "C" int creste_conv2d_wgrad_tc(const creste_conv_desc* d, const float* x, const float* g, float* dw, void* ws,
                                      size_t ws_bytes, void* stream) {
  if (!d || !x || !g || !dw) { creste::set_error("creste_conv2d_wgrad_tc: bad args"); return CRESTE_ERR_ARG; }
  return creste::wgrad_tc_launch(d, x, g, dw, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int creste_conv2d_wgrad_tc_presplit(const creste_conv_desc* d, const void* x_hi, const void* x_lo,
                                               const float* x_scal, const void* g_hi, const void* g_lo,
                                               const float* g_scal, float* dw, void* ws, size_t ws_bytes, void* stream) {
  if (!d || !x_hi || !x_lo || !x_scal || !g_hi || !g_lo || !g_scal || !dw) {
    creste::set_error("creste_conv2d_wgrad_tc_presplit: bad args");
    return CRESTE_ERR_ARG;
  }
  return creste::wgrad_tc_presplit_launch(d, x_hi, x_lo, x_scal, g_hi, g_lo, g_scal, dw, ws, ws_bytes, (cudaStream_t)stream);
}
extern "C" int creste_f16_split(const float* x, long long numel, void* hi, void* lo, float* scal, void* stream) {
  if (!x || !hi || !lo || !scal || numel <= 0 || (numel & 7)) { creste::set_error("creste_f16_split: bad args (numel %% 8 == 0)"); return CRESTE_ERR_ARG; }
  return creste::f16_split_launch(x, (size_t)numel, hi, lo, scal, (cudaStream_t)stream);
}
extern "C" int creste_f16_split_amax(const float* x, long long numel, const float* amax, void* hi, void* lo, float* scal,
                                     void* stream) {
  if (!x || !amax || !hi || !lo || !scal || numel <= 0 || (numel & 7)) { creste::set_error("creste_f16_split_amax: bad args (numel %% 8 == 0)"); return CRESTE_ERR_ARG; }
  return creste::f16_split_known_launch(x, (size_t)numel, hi, lo, scal, amax, (cudaStream_t)stream);
}

/* 3xFP16 weight operand of creste_conv2d (precision 4) in one launch: logical w[k][c][r][s] read through element
 * strides; out = npad*R*S*cpad64 fp16 hi, the same count of lo, then npad fp32 inverse scales
 * (creste_conv2d_tc_layout gives npad; cpad64 = C rounded up to 64). */
extern "C" int creste_pack_weight_f16(const float* w, long long sK, long long sC, long long sR, long long sS, int K,
                                      int C, int R, int S, float* out, void* stream) {
  if (!w || !out || K <= 0 || C <= 0 || R <= 0 || S <= 0) { creste::set_error("creste_pack_weight_f16: bad args"); return CRESTE_ERR_ARG; }
  return creste::pack_weight_f16_launch(w, sK, sC, sR, sS, K, C, R, S, out, (cudaStream_t)stream);
}
