/*
 * creste_b200.h -- C ABI of libcreste_b200.so: the sm_100a kernels behind the CREStE
 * perception->costmap + IRL hot path.
 *
 * The reference (ut-amrl/creste_public) is pure Python/PyTorch and has no FFI layer; each entry
 * point below replaces a group of PyTorch library dispatches inside one reference function
 * (file:line cited per entry; paths relative to the reference root).  The reference-side
 * binding is a ctypes stub inside the nn.Module that owns the op -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed for the duration of the call (PyTorch owns all
 *     memory); nothing is retained or freed; scratch comes from the caller (`ws`, `ws_bytes`);
 *   - all tensors are dense, fp32 unless stated; activations of the conv family are NHWC;
 *   - `stream` is a cudaStream_t passed as void* (use torch.cuda.current_stream().cuda_stream);
 *     no call synchronises the device unless documented;
 *   - return value: 0 = ok, otherwise a negative creste error or a positive cudaError_t;
 *     creste_last_error() returns a thread-local message for the last failure.
 */
#ifndef CRESTE_B200_H
#define CRESTE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRESTE_OK 0
#define CRESTE_ERR_ARG (-1)       /* bad argument / unsupported shape */
#define CRESTE_ERR_WORKSPACE (-2) /* workspace too small */
#define CRESTE_ERR_NO_DEVICE (-3) /* no sm_100 device */

int creste_version(void);
const char* creste_last_error(void);
/* number of SMs of the current device (0 when no device) */
int creste_num_sms(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long creste_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Value iteration.  Replaces VIN.value_iteration_manual, creste/models/blocks/vin.py:48-80
 * (stencil weights vin.py:36-46): v<-0; repeat { q = conv3x3(r + gamma*v); v' = max_a q;
 * delta = max over the WHOLE batch |v'-v| } while delta > thr; then q = conv(r+gamma*v),
 * pi = softmax_a(q).  One persistent cooperative launch, device-side convergence test.
 *   r [B,H,W]; v_out [B,H,W]; q_out, pi_out [B,8,H,W] (either may be NULL);
 *   sweeps_out: DEVICE int[2] = {number of sweeps K, 1 if max_sweeps was hit};
 *   ws: >= creste_vi_workspace_bytes(B,H,W,max_sweeps) bytes. */
size_t creste_vi_workspace_bytes(int B, int H, int W, int max_sweeps);
int creste_vi_solve(const float* r, float* v_out, float* q_out, float* pi_out, int B, int H,
                    int W, float gamma, float thr, int max_sweeps, int* sweeps_out, void* ws,
                    size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Expected state-visitation frequency + greedy rollout.  Replaces
 * MaxEntIRL.expected_state_visitation_frequency, creste/models/lfd.py:156-277 and
 * earliest_pose_in_fov, creste/utils/train_utils.py:765-803.
 *   policy [B,8,H,W]; expert_rc [B,T_expert,2] = expert[:,:,:2,2] (row, col; un-pooled BEV
 *   cells); T = action_horizon; fov [H,W] uint8; ds = reward_cfg.ds; sharpen/temperature =
 *   policy_kwargs;
 *   exp_svf [B,H,W]; states [B,T,2] int64; states_grid [B,H,W];
 *   ws: >= creste_svf_workspace_bytes(B,H,W,T). */
size_t creste_svf_workspace_bytes(int B, int H, int W, int T);
int creste_svf(const float* policy, const float* expert_rc, const uint8_t* fov, int B, int H, int W,
               int T, int T_expert, int ds, int sharpen, float temperature, int zero_terminal_state,
               float* exp_svf, int64_t* states, float* states_grid, void* ws, size_t ws_bytes,
               void* stream);

/* ---------------------------------------------------------------------------------------------
 * Camera frustum -> LiDAR xyz -> BEV voxel coordinates.  Replaces Camera2World.forward
 * (creste/models/blocks/splat_projection.py:19-51), the bounds mask (:169) and
 * _points_to_voxels (:175-189).
 *   depth [N,Hs,Ws] metres; p2p [N,4,4]; range[6] = point_cloud_range; voxel[2] (host floats);
 *   xy [N,P,2] float voxel coords (the reference's `bev_coords`), z [N,P], mask [N,P] uint8. */
int creste_frustum_to_bev(const float* depth, const float* p2p, int N, int Hs, int Ws,
                          const float* range_host6, const float* voxel_host2, float* xy, float* z,
                          uint8_t* mask, void* stream);

/* z -> MLP(1->64->32, ReLU) and concat with the 256 image features into one NHWC row per point.
 * Replaces splat_projection.py:152-158 (z_proj) + the torch.cat at :158.
 *   feats NHWC [NP, C]; z [NP]; w1[64], b1[64], w2[32,64], b2[32]; out NHWC [NP, C+32]. */
int creste_zmlp_concat(const float* feats, const float* z, int NP, int C, const float* w1,
                       const float* b1, const float* w2, const float* b2, float* out, void* stream);
/* the same with max|out| published to amax_out (DEVICE float[1] or NULL) for the fusion conv's operand scale */
int creste_zmlp_concat_ex(const float* feats, const float* z, int NP, int C, const float* w1, const float* b1,
                          const float* w2, const float* b2, float* out, float* amax_out, void* stream);

/* Camera2World.forward as a stand-alone op, creste/models/blocks/splat_projection.py:19-51:
 * xyz[n, :, v, u] = (p2p[n] @ [u*d, v*d, d, 1])[:3] (K = 4 in-order FMA chain, the CPU bmm order).
 *   depth [N,Hs,Ws]; p2p [N,4,4]; xyz [N,3,Hs,Ws].  The fused path (creste_frustum_to_bev) never
 *   materialises xyz; this entry exists for callers of the reference's Camera2World module. */
int creste_camera_to_world(const float* depth, const float* p2p, int N, int Hs, int Ws, float* xyz,
                           void* stream);
/* Camera2MapMulti._points_to_voxels, splat_projection.py:175-189:
 * xy = (lidar2map @ [x,y,z,1])[:2] / voxel_size[:2].
 *   pts [NP,3] device; lidar2map HOST float[16] row-major; voxel HOST float[2]; xy [NP,2] device. */
int creste_points_to_voxels(const float* pts, long long NP, const float* lidar2map, const float* voxel,
                            float* xy, void* stream);

/* Bilinear 4-tap scatter-add splat with mean normalisation.  Replaces
 * Camera2MapMulti.splat_soft, creste/models/blocks/splat_projection.py:262-354
 * (scatter_mode='mean') and the `feats * xyz_mask` at :219.
 *   xy [N,P,2]; feats NHWC [N,P,F]; mask [N,P] uint8 (NULL = all ones);
 *   bev_nhwc [N,H,W,F] (may be NULL), bev_nchw [N,F,H,W] (may be NULL), dens [N,H,W];
 *   idx_out [N,P,4] int64 (may be NULL): linear cell index of each tap in the reference's
 *   (xdiff,ydiff) loop order, -1 when the tap is out of bounds -- the bit-exact criterion.
 *   ws: >= N*H*W*F*4 bytes (accumulator) */
size_t creste_splat_workspace_bytes(int N, int H, int W, int F);
int creste_splat_soft(const float* xy, const float* feats, const uint8_t* mask, int N, int P, int F,
                      int H, int W, float min_weight, float* bev_nhwc, float* bev_nchw, float* dens,
                      int64_t* idx_out, void* ws, size_t ws_bytes, void* stream);
/* Backward of creste_splat_soft (autograd of Camera2MapMulti.splat_soft, splat_projection.py:262-354, which the
 * reference's stage-2 training step differentiates into the point features AND the voxel coordinates).
 *   xy [N,P,2], feats NHWC [N,P,F], mask [N,P] (NULL = all valid): the forward inputs;
 *   bev_nhwc [N,H,W,F], dens [N,H*W]: the forward outputs; g_bev_nhwc / g_dens (NULL = 0): their gradients;
 *   dfeats [N,P,F] (0 for masked points), dxy [N,P,2] (floor has no gradient);
 *   ws: >= creste_splat_bwd_workspace_bytes(N,H,W). */
size_t creste_splat_bwd_workspace_bytes(int N, int H, int W);
int creste_splat_soft_bwd(const float* xy, const float* feats, const uint8_t* mask, const float* bev_nhwc,
                          const float* dens, const float* g_bev_nhwc, const float* g_dens, int N, int P, int F,
                          int H, int W, float min_weight, float* dfeats, float* dxy, void* ws, size_t ws_bytes,
                          void* stream);
/* Backward of creste_frustum_to_bev w.r.t. the depth (xy and z are affine in it): autograd of
 * Camera2World.forward + _points_to_voxels (splat_projection.py:19-51, :175-189).
 *   dxy [N,P,2] and / or dz [N,P] (either may be NULL); voxel HOST float[2]; ddepth [N,Hs,Ws]. */
int creste_frustum_bwd(const float* dxy, const float* dz, const float* p2p, int N, int Hs, int Ws,
                       const float* voxel, float* ddepth, void* stream);


/* ---------------------------------------------------------------------------------------------
 * LiDAR -> sparse depth raster.  Replaces pixels_to_depth, creste/utils/projection.py:64-134
 * (float64 projection, truncation to int32, per-pixel max) and the mm quantisation of
 * scripts/preprocessing/build_dense_depth.py:461-463.
 *   pc [npts, stride] float32 (xyz first); P34 host double[12] (lidar2camrect);
 *   depth_m [H,W] float32 (may be NULL), depth_mm [H,W] float32 (may be NULL);
 *   ws: >= H*W*8 bytes. */
int creste_lidar_raster(const float* pc, int npts, int stride, const double* P34_host, int H, int W,
                        float* depth_m, float* depth_mm, void* ws, size_t ws_bytes, void* stream);

/* Softmax-expectation metric depth + arg-max bin.  Replaces
 * convert_to_metric_depth_differentiable, creste/utils/depth_utils.py:300-313 and
 * DepthCompletion._convert_to_metric_depth, creste/models/depth.py:60-100.
 *   logits NHWC [NP, D] (D = 128); metric [NP] = expectation / out_div (out_div = 1000: metres as
 *   depth.py:100 returns; 1: the units of depth_min / depth_max as depth_utils.py:300-313 returns);
 *   bins [NP] int64 (may be NULL). */
int creste_depth_expectation(const float* logits, int NP, int D, float depth_min_mm,
                             float depth_max_mm, float out_div, float* metric, int64_t* bins,
                             void* stream);

/* bin_depths, creste/utils/depth_utils.py:346-383.  mode 0 = UD, 1 = LID, 2 = SID.
 *   depth [n]; out_f [n] float indices (target == 0) or out_i [n] int64 with out-of-range / non-finite
 *   indices set to num_bins (target != 0).  Exactly one of out_f / out_i is written. */
int creste_bin_depths(const float* depth, long long n, int mode, float depth_min, float depth_max,
                      int num_bins, int target, float* out_f, int64_t* out_i, void* stream);
/* Backward of creste_depth_expectation w.r.t. the logits: d logit_k = g * softmax_k * (val_k - E) / out_div
 * (autograd of convert_to_metric_depth_differentiable, depth_utils.py:300-313).  logits NHWC [NP,128]. */
int creste_depth_expectation_bwd(const float* logits, const float* g_metric, int NP, int D, float depth_min_mm,
                                 float depth_max_mm, float out_div, float* dlogits, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Convolution family (NHWC, fp32 storage).  One descriptor covers every dense conv / linear on
 * the path: F.conv2d + BatchNorm(eval) + activation (+ residual) as issued
 * by creste/models/blocks/effnet.py:12-28,74; conv.py:22-29,48-55,63-85; inpainting.py:52-68,
 * 96-109; torchvision BasicBlock; efficientnet_pytorch MBConv 1x1 convs. */
typedef struct {
  int N, H, W, C;       /* input NHWC */
  int K;                /* output channels */
  int R, S;             /* filter height, width */
  int stride;
  int pad_t, pad_l;     /* zero padding on top / left (bottom/right implied by P, Q) */
  int P, Q;             /* output height, width */
  int act;              /* 0 none, 1 relu, 2 swish, 3 sigmoid */
  int out_nchw;         /* 0: out is NHWC [N,P,Q,K]; 1: out is NCHW [N,K,P,Q] */
  int precision;        /* 0: fp32 FFMA (exact fp32 products, CUDA cores);
                           1: 3xTF32 split on tcgen05 (fp32-faithful);
                           2: single-pass TF32 on tcgen05; 3: reserved;
                           4: 3xFP16 split on tcgen05 kind::f16 (fp32-faithful, default): fp16
                              hi/lo operands under exact power-of-two scales (per tensor for the
                              activations -- computed on the device --, per output channel for
                              the weights) */
} creste_conv_desc;

/* w_packed: [R*S*C, K] row-major (k index = (r*S+s)*C + c); scale/shift [K] (folded BN / bias;
 * NULL = 1 / 0); gate [N,C] multiplies the input per (n,c) (SE gate; NULL = none);
 * residual NHWC [N,P,Q,K] added before the activation (NULL = none). */
int creste_conv2d(const creste_conv_desc* d, const float* x, const float* w_packed,
                  const float* scale, const float* shift, const float* gate, const float* residual,
                  float* out, void* ws, size_t ws_bytes, void* stream);
/* creste_conv2d with the operand-scale bound carried beside the tensors (3xFP16 mode):
 *   amax_in  DEVICE float[1] or NULL: an upper bound of max|x * gate| (e.g. the amax_out of the producing call);
 *            given, the operand pre-pass skips its amax pass over x and derives the power-of-two scale from it;
 *   amax_out DEVICE float[1] or NULL: receives max|out| (zeroed, then atomicMax from the epilogue). */
int creste_conv2d_ex(const creste_conv_desc* d, const float* x, const float* w_packed, const float* scale,
                     const float* shift, const float* gate, const float* residual, float* out,
                     const float* amax_in, float* amax_out, void* ws, size_t ws_bytes, void* stream);
/* Tensor-core conv on operands that are ALREADY split (precision 4 = 3xFP16, 5 = single-pass fp16): x_hi / x_lo are
 * dense fp16 [N,H,W,C] tensors holding fp16(x*s) and fp16((x*s - hi) * 2^11), x_scal = DEVICE float[2] {s, 1/s} --
 * written by creste_upsample_concat_split (or any producer that knows a bound of its output).  No pre-pass over the
 * activations; no gate operand.  x_lo may be NULL for precision 5. */
int creste_conv2d_presplit(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                           const float* w_packed, const float* scale, const float* shift, const float* residual,
                           float* out, float* amax_out, void* stream);
/* The two calls above with the output ALSO (out != NULL) or ONLY (out == NULL) written as the next tensor-core conv's
 * 3xFP16 operand: out_hi / out_lo fp16 [N,P,Q,K] (out_lo may be NULL for precision 5), out_scal = DEVICE float[2]
 * {s, 1/s}.  The power-of-two scale comes from an a-priori bound of max|out| the caller supplies as two host floats:
 *   bound_mul = max_k(sum_{c,r,s} |w[k,c,r,s]| * |scale[k]|),  bound_add = max_k |shift[k]|
 * (|out| <= max|x| * bound_mul + bound_add; residual must be NULL).  max|x| is read on the device: amax_in / x_amax
 * (DEVICE float[1], the TRUE maximum carried with the input -- chained a-priori bounds would compound) when given, else
 * 2^15 / s_in from the input's scale record.  Replaces the split pre-pass of the consuming conv
 * in conv -> BN -> ReLU -> conv chains (reference creste/models/blocks/effnet.py:12-28, inpainting.py:52-68 and the
 * torchvision BasicBlocks of :80-103).  Precision 4 / 5, NHWC output, K % 8 == 0. */
int creste_conv2d_split_out(const creste_conv_desc* d, const float* x, const float* w_packed, const float* scale,
                            const float* shift, const float* gate, const float* residual, float* out,
                            const float* amax_in, float* amax_out, void* out_hi, void* out_lo, float* out_scal,
                            float bound_mul, float bound_add, void* ws, size_t ws_bytes, void* stream);
int creste_conv2d_presplit_split_out(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                                     const float* w_packed, const float* scale, const float* shift,
                                     const float* residual, float* out, float* amax_out, void* out_hi, void* out_lo,
                                     float* out_scal, float bound_mul, float bound_add, const float* x_amax,
                                     void* stream);

size_t creste_conv2d_workspace_bytes(const creste_conv_desc* d);
/* tcgen05 path (precision 1, 2): 1 if the shape is served by the tensor-core kernel (stride 1 or 2, R, S <= 7,
 * C % 4 == 0, K >= 8, >= 128 output pixels), else the caller must use precision 0.  For those
 * modes w_packed is [Npad][R*S*Cpad] fp32 pre-rounded to tf32 ("hi"), followed for precision 1
 * by the same-shaped "lo" = rna_tf32(w - hi); creste_conv2d_tc_layout reports Npad (K rounded up
 * to a multiple of the N tile block_n) and Cpad (C rounded up to 32). */
int creste_conv2d_tc_supported(const creste_conv_desc* d);
/* development aid (tools/conv_timeline.py): while dev_buf != NULL every tensor-core conv CTA writes
 * 6 globaltimer stamps (start, prologue done, first operands, main loop done, epilogue done, end)
 * to dev_buf[blockIdx.x * 8 ...] (uint64); pass NULL to switch it off. */
int creste_conv2d_tc_debug(void* dev_buf);
int creste_conv2d_tc_layout(int K, int C, int R, int S, int* block_n, int* npad, int* cpad);

/* Depthwise conv + folded BN + swish, also producing the per-(n,c) spatial partial sums that the
 * SE block needs (efficientnet_pytorch MBConvBlock: _depthwise_conv -> _bn1 -> swish -> avg-pool).
 * Partial sums are written per pixel tile in a fixed order (no atomics: bit-reproducible).
 *   x NHWC [N,H,W,C]; w [R*S, C]; out NHWC [N,P,Q,C];
 *   chan_part [N, nparts, C] with nparts = creste_dwconv_num_parts(N, P, Q). */
int creste_dwconv_num_parts(int N, int P, int Q);
/* Row count of chan_part for the TILED depthwise kernel (one row per 8 x 32 / 8 x 16 output tile, stride 1 / 2): a
 * caller that allocates chan_part with this many rows (and passes it as nparts) gets the shared-memory-tiled kernel,
 * creste_dwconv_num_parts rows the x-blocked one.  Same results per output element; the squeeze-excite partial sums
 * are grouped by tile instead of by pixel run. */
int creste_dwconv_tile_parts(int P, int Q, int stride);
/* the row count of whichever of the two kernels is faster for this layer (what the mirror allocates) */
int creste_dwconv_parts(int N, int P, int Q, int R, int stride);
int creste_dwconv_bn_swish(const float* x, const float* w, const float* scale, const float* shift,
                           int N, int H, int W, int C, int R, int stride, int pad_t, int pad_l,
                           int P, int Q, float* out, float* chan_part, int nparts, void* stream);
/* the same with max|out| published to amax_out (DEVICE float[1], zeroed first; may be NULL): the bound travels with
 * the tensor so that the project conv's 3xFP16 operand pre-pass needs no amax pass (the SE gate is a sigmoid:
 * max|out * gate| <= max|out|).  Stride 1 or 2. */
int creste_dwconv_bn_swish_ex(const float* x, const float* w, const float* scale, const float* shift, int N, int H,
                              int W, int C, int R, int stride, int pad_t, int pad_l, int P, int Q, float* out,
                              float* chan_part, int nparts, float* amax_out, void* stream);

/* SE gate: mean -> 1x1 reduce(+b) -> swish -> 1x1 expand(+b) -> sigmoid.
 *   chan_part [N,nparts,C] (summed in order); w_red [Csq,C], b_red [Csq], w_exp [C,Csq],
 *   b_exp [C]; gate [N,C]. */
int creste_se_gate(const float* chan_part, int nparts, float inv_hw, int N, int C, int Csq,
                   const float* w_red,
                   const float* b_red, const float* w_exp, const float* b_exp, float* gate,
                   void* stream);

/* cat([skip, bilinear_upsample(x, align_corners=False)], channel) in NHWC.  Replaces nn.Upsample +
 * torch.cat in Up.forward (creste/models/blocks/effnet.py:26-28), DeconvHead.up2[0]
 * (inpainting.py:56) and MultiScaleFCN trunk upsample (conv.py:128).
 *   skip NHWC [N,Ho,Wo,Cs] (NULL / Cs = 0: no concat); x NHWC [N,Hi,Wi,Cx];
 *   rh, rw: source-per-destination ratios (1/scale_factor); out NHWC [N,Ho,Wo,Cs+Cx];
 *   x_first = 0: channels [skip, up(x)] (Up.forward); 1: [up(x), skip] (MultiScaleFCN.forward,
 *   conv.py:156). */
int creste_upsample_concat(const float* skip, int Cs, const float* x, int N, int Hi, int Wi, int Cx,
                           int Ho, int Wo, float rh, float rw, int x_first, float* out,
                           void* stream);
/* creste_upsample_concat writing the 3xFP16 operand of the consuming conv directly (hi / lo fp16 [N,Ho,Wo,Cs+Cx] and
 * scal = {s, 1/s}); amax_a / amax_b: DEVICE float[1] bounds of max|skip| and max|x| (amax_b may be NULL when there is
 * no skip).  lo may be NULL (single-pass fp16 consumer).  (Cs + Cx) % 8 == 0. */
int creste_upsample_concat_split(const float* skip, int Cs, const float* x, int N, int Hi, int Wi, int Cx, int Ho,
                                 int Wo, float rh, float rw, int x_first, const float* amax_a, const float* amax_b,
                                 void* hi, void* lo, float* scal, void* stream);


/* 2x2/2 max-pool over the channel-concatenation of up to 3 NCHW or NHWC sources, cropped to the
 * first `rows_out` output rows.  Replaces vin.py:104-115 (cat + max_pool2d + crop) and
 * conv.py:117 (trunk MaxPool2d).   srcs: `nsrc` NHWC tensors [N,H,W,Ci]; out NHWC
 * [N,rows_out,W/2,sum Ci]; out_nchw (may be NULL) [N,sum Ci,rows_out,W/2]. */
int creste_maxpool2_concat(const float* const* srcs_host, const int* chans_host, int nsrc, int N,
                           int H, int W, int rows_out, float* out_nhwc, float* out_nchw,
                           void* stream);

/* layout shuffles used at the module boundary (the reference's tensors are NCHW) */
int creste_nchw_to_nhwc(const float* in, int N, int C, int H, int W, float* out, void* stream);
int creste_nhwc_to_nchw(const float* in, int N, int H, int W, int C, float* out, void* stream);
/* Projection head: 1x1 conv C -> K (K <= 32; the DeconvHead `proj`, creste/models/blocks/inpainting.py:52-68)
 * fused with the NHWC -> NCHW change of its input (the `{prefix}_features` entry of the output dict).
 *   x NHWC [N,H,W,C] (C % 32 == 0); w [K,C]; bias [K] or NULL; pred_nhwc [N,H,W,K] / pred_nchw [N,K,H,W] (either may
 *   be NULL); x_nchw [N,C,H,W] or NULL.  Exact fp32 FFMA chain over ascending channels. */
int creste_proj_head(const float* x, const float* w, const float* bias, int N, int H, int W, int C, int K,
                     float* pred_nhwc, float* pred_nchw, float* x_nchw, void* stream);
/* Layout helpers of the strided-convolution gradients (the stride-2 7x7 / 3x3 / 1x1 convs of the ResNet-18 BEV
 * trunk, creste/models/blocks/inpainting.py:80-90, differentiated by the reference's stage-2 step):
 *   creste_dilate       z [N,Hz,Wz,C] = 0 except z[n, p*stride, q*stride, :] = g[n,p,q,:]   (data gradient =
 *                       stride-1 conv of the dilated output gradient with the flipped weights)
 *   creste_phase_slice  out [N,Ha,Wa,C] = x[n, a + stride*i, b + stride*j, :], zero outside the image (weight
 *                       gradient = stride-1 weight gradients over the stride^2 phase images).  C % 4 == 0. */
int creste_dilate(const float* g, int N, int P, int Q, int C, int stride, int Hz, int Wz, float* z, void* stream);
int creste_phase_slice(const float* x, int N, int H, int W, int C, int stride, int a, int b, int Ha, int Wa,
                       float* out, void* stream);

/* Expert / counterfactual visitation raster.  Replaces MaxEntIRLLoss.compute_expert_visitation,
 * creste/utils/loss_utils.py:1055-1116 (second definition).
 *   traj [B,T,2] (row,col) float32 (is_f64 = 0) or float64 (is_f64 = 1), un-pooled cells;
 *   max_steps: host int = ceil(max segment length) (the reference's `.item()` sync);
 *   counts [B,H,W] in {0,1}. */
int creste_expert_visitation(const void* traj, int is_f64, int B, int T, double map_ds,
                             int max_steps, int H, int W, float* counts, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-3 (MaxEnt / counterfactual IRL) training step.  The reference trains the reward FCN
 * (creste/models/blocks/conv.py:88-161, train-mode BatchNorm) through PyTorch autograd, with a
 * double backward for the SMODICE gradient penalty (creste/utils/loss_utils.py:1208-1217), inside
 * MaxEntIRLModel.training_step (creste/train_traversability.py:62-103).  The entries below are the
 * primitives of that graph and of the graph of its backward (a set closed under differentiation;
 * creste_public_b200/autograd.py composes them).  Channels-last fp32 throughout. */

/* y[pix,c] = act(x[pix,c]*a[c] + b[c]); a/b NULL = 1/0; relu: 0/1.  Replaces the elementwise part
 * of F.batch_norm (train) and nn.ReLU in conv.py:63-85,117-126. */
int creste_chan_affine(const float* x, const float* a, const float* b, long long npix, int C,
                       int relu, float* y, void* stream);
/* out = g * (y > 0): ReLU backward (autograd's threshold_backward). */
int creste_relu_bwd(const float* g, const float* y, long long n, float* out, void* stream);
/* the two above, also publishing max|out| (DEVICE float[1], zeroed here) for creste_f16_split_amax: in the stage-3
 * graph (conv.py:117-126 differentiated twice) these are the producers of most tensor-core conv operands */
int creste_chan_affine_amax(const float* x, const float* a, const float* b, long long npix, int C,
                            int relu, float* y, float* amax_out, void* stream);
int creste_relu_bwd_amax(const float* g, const float* y, long long n, float* out, float* amax_out, void* stream);
/* out[c] = sum_pix x[pix,c] * (y ? y[pix,c] : 1): BatchNorm batch statistics and the channel
 * reductions of its backward; two-stage fixed-order reduction.  ws >= the _workspace_bytes. */
size_t creste_chan_dot_workspace_bytes(long long npix, int C);
int creste_chan_dot(const float* x, const float* y, long long npix, int C, float* out, void* ws,
                    size_t ws_bytes, void* stream);
/* out2 DEVICE double[2*C] = {sum_pix x}, {sum_pix x^2}: F.batch_norm's batch statistics in one pass.
 * ws >= 2 * creste_chan_dot_workspace_bytes(npix, C). */
int creste_chan_stats(const float* x, long long npix, int C, double* out2, void* ws, size_t ws_bytes,
                      void* stream);
/* 2x2/2 max-pool backward (dx[argmax] = g; first maximum wins, the PyTorch tie rule) and its
 * adjoint (out[pooled] = gg[argmax]) -- conv.py:117 under autograd.  x [N,H,W,C]. */
int creste_maxpool2_bwd(const float* x, const float* g, int N, int H, int W, int C, float* dx,
                        void* stream);
int creste_maxpool2_gather(const float* x, const float* gg, int N, int H, int W, int C, float* out,
                           void* stream);
/* adjoint of the bilinear upsample of creste_upsample_concat (conv.py:128 under autograd):
 * g [N,Ho,Wo,C] -> dx [N,Hi,Wi,C]; rh, rw = source-per-destination ratios. */
int creste_upsample_adjoint(const float* g, int N, int Hi, int Wi, int C, int Ho, int Wo, float rh,
                            float rw, float* dx, void* stream);
/* the same reading the channel slice [c0, c0 + C) of g [N,Ho,Wo,Cg] in place: backward of `cat([x2, up(x1)])`
 * (effnet.py:22-28, inpainting.py:56) without a copy of the up-sampled part */
int creste_upsample_adjoint_slice(const float* g, int Cg, int c0, int N, int Hi, int Wi, int C, int Ho, int Wo,
                                  float rh, float rw, float* dx, void* stream);
/* weight gradient of a stride-1 conv (autograd's convolution_backward, weight part):
 * x [N,H,W,C], g [N,P,Q,K] -> dw_packed [R*S*C, K]; C, K multiples of 8, <= 64. */
size_t creste_conv2d_wgrad_workspace_bytes(const creste_conv_desc* d);
int creste_conv2d_wgrad(const creste_conv_desc* d, const float* x, const float* g, float* dw_packed,
                        void* ws, size_t ws_bytes, void* stream);
/* per-sample reductions / scalings of the loss (loss_utils.py:1142-1146, 1195-1203):
 * row_dot: out[b] = sum_i x[b,i]*y[b,i]*mask[b,i] (y, mask may be NULL);
 * row_scale: out[b,i] = x[b,i]*s[b]*mask[b,i]; row_normalize: x*mask / (sum(x*mask) + eps). */
int creste_row_dot(const float* x, const float* y, const uint8_t* mask, int B, long long n,
                   float* out, void* stream);
int creste_row_scale(const float* x, const float* s, const uint8_t* mask, int B, long long n,
                     float* out, void* stream);
int creste_row_normalize(const float* x, const uint8_t* mask, int B, long long n, float eps,
                         float* out, void* stream);
/* SMODICE gradient penalty (loss_utils.py:1216-1217): G NCHW [B,C,HW];
 * penalty = mean_{b,pixel} (||G[b,:,pixel]||_2 - 1)^2 (device scalar); _bwd: dG = g_scalar * dP/dG. */
size_t creste_grad_penalty_workspace_bytes(int B, long long HW);
int creste_grad_penalty(const float* G, int B, int C, long long HW, float* penalty_out, void* ws,
                        size_t ws_bytes, void* stream);
int creste_grad_penalty_bwd(const float* G, const float* g_scalar, int B, int C, long long HW,
                            float* dG, void* stream);
/* torch.optim.Adam step (train_traversability.py optimizer; no amsgrad / weight decay) on flat
 * buffers; g is multiplied by grad_scale first (1/world after the NCCL sum all-reduce). */
int creste_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1,
                     float b2, float eps, int step, float grad_scale, void* stream);

/* Stage-1 loss VALUES (forward; the gradients are creste_ce_depth_bwd / creste_masked_mse_bwd below):
 * CrossEntropyDepth + SmoothL1Depth (creste/utils/loss_utils.py:477-573 with bin_depths mode "UD",
 * creste/utils/depth_utils.py:346-383) in one pass over the NCHW depth logits.
 *   logits [N,D,HW]; pred_bins int64 [N,HW] (depth_preds_bins: what the shipped config feeds the
 *   Smooth-L1 term); label_mm [N,HW]; acc4 DEVICE double[4] = {sum CE over valid pixels, #valid,
 *   #(argmax == gt bin), sum smooth_l1(pred_bins - label/1000)}. */
int creste_stage1_depth_losses(const float* logits_nchw, const long long* pred_bins,
                               const float* label_mm, int N, int D, long long HW, float depth_min,
                               float depth_max, float beta, double* acc4, void* stream);
/* MSELoss on the DINO feature targets (loss_utils.py:606-647, overlap_only = False):
 * acc2 DEVICE double[2] = {sum (pred-gt)^2 over elements with !isinf(gt), their count}. */
int creste_masked_mse(const float* pred, const float* gt, long long n, double* acc2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-1 (distillation) backbone TRAINING step: the train-mode graph of DistillationBackbone
 * (creste/models/distillation.py:145-207; EfficientNet-B0 trunk + U-Net decoder,
 * creste/models/blocks/effnet.py:8-98; heads creste/models/blocks/conv.py:5-32) and its backward,
 * as driven by creste/train_pefree.py:76-106.  Dense convs reuse creste_conv2d /
 * creste_conv2d_wgrad; everything else is below (csrc/backbone_train.cu).  NHWC fp32, C % 4 == 0.
 * act codes: 0 none, 1 relu, 2 swish (x*sigmoid(x)), 3 sigmoid. */
/* workspace of the two-stage channel reductions: nacc terms, Z samples (Z = 1: whole batch) */
size_t creste_chan_reduce_workspace_bytes(long long npix, int C, int nacc, int Z);
/* BatchNorm batch statistics (F.batch_norm training=True): out2 DEVICE double[2*C] = {sum x},
 * {sum x^2}; any C % 4 == 0 (the EfficientNet mid tensors reach 1152 channels). */
int creste_chan_moments(const float* x, long long npix, int C, double* out2, void* ws, size_t ws_bytes,
                        void* stream);
/* per-channel algebra of BatchNorm(training=True) in one launch each.  _fwd_: moments -> ab [2*C] float = (a, b) of
 * y = x*a + b, mean_inv [2*C] double = (mean, 1/sqrt(var+eps)), and -- when running_mean/var are given -- the
 * momentum update with the unbiased variance.  _bwd_: sums2 = (sum gu, sum gu*x) -> out4 [4*C] float =
 * (dgamma, dbeta, q, r) with dx = gu*a + x*q + r. */
int creste_bn_fwd_finalize(const double* stats2, const float* weight, const float* bias, int C, double M, double eps,
                           float momentum, float* running_mean, float* running_var, float* ab, double* mean_inv,
                           void* stream);
int creste_bn_bwd_finalize(const double* sums2, const float* ab, const double* mean_inv, int C, double M, float* out4,
                           void* stream);
/* y = act(x*a[c] + b[c])  (BatchNorm affine + ReLU / swish in one pass) */
int creste_chan_affine_act(const float* x, const float* a, const float* b, long long npix, int C, int act,
                           float* y, void* stream);
/* the same, also publishing max|y| (DEVICE float[1], zeroed here) for the 3xFP16 operand scale of the conv that
 * consumes y (creste_f16_split_amax): the consumer makes no amax pass.  Same reduction as the split pre-pass, same
 * bits.  Replaces nothing in the reference (torch's cuDNN path has no operand scale); it is part of
 * F.batch_norm + activation, creste/models/blocks/effnet.py:14-20 / efficientnet_pytorch MBConvBlock. */
int creste_chan_affine_act_amax(const float* x, const float* a, const float* b, long long npix, int C, int act,
                                float* y, float* amax_out, void* stream);
/* first half of the BatchNorm(+act) backward: gu = g * act'(x*a[c]+b[c]) (written when act != 0) and
 * sums2 DEVICE double[2*C] = {sum gu}, {sum gu*x}; ws >= creste_chan_reduce_workspace_bytes(npix,C,2,1);
 * gu may be NULL (not stored: creste_chan_axpby_act recomputes it) */
int creste_bn_act_bwd(const float* g, const float* x, const float* a, const float* b, long long npix, int C,
                      int act, float* gu, double* sums2, void* ws, size_t ws_bytes, void* stream);
/* second half: out = u*p[c] + x*q[c] + r[c] */
int creste_chan_axpby(const float* u, const float* x, const float* p, const float* q, const float* r,
                      long long npix, int C, float* out, void* stream);
/* the same, also publishing max|out| (DEVICE float[1], zeroed here): out is the gradient the data- / weight-gradient
 * convs of the layer below split into their 3xFP16 operand */
int creste_chan_axpby_amax(const float* u, const float* x, const float* p, const float* q, const float* r,
                           long long npix, int C, float* out, float* amax_out, void* stream);
/* second half with gu recomputed: out = (g * act'(x*a[c]+b[c])) * p[c] + x*q[c] + r[c]  (creste_bn_act_bwd may then be
 * called with gu = NULL); amax_out optional (DEVICE float[1], zeroed here).  Same bits as the two-kernel form. */
int creste_chan_axpby_act(const float* g, const float* x, const float* a, const float* b, int act, const float* p,
                          const float* q, const float* r, long long npix, int C, float* out, float* amax_out,
                          void* stream);
/* depthwise R x R conv (R in {3,5}, stride in {1,2}) with the static TF-'SAME' padding of
 * efficientnet_pytorch (low pads given, high implied by P, Q): w [R*R][C]; x [N,H,W,C]; y [N,P,Q,C];
 * _dgrad: dx from g [N,P,Q,C]; _wgrad: dw [R*R][C]. */
int creste_dwconv_fwd(const float* x, const float* w, int N, int H, int W, int C, int R, int stride,
                      int pad_t, int pad_l, int P, int Q, float* y, void* stream);
int creste_dwconv_dgrad(const float* g, const float* w, int N, int H, int W, int C, int R, int stride,
                        int pad_t, int pad_l, int P, int Q, float* dx, void* stream);
size_t creste_dwconv_wgrad_workspace_bytes(int N, int C, int R, int P, int Q);
int creste_dwconv_wgrad(const float* x, const float* g, int N, int H, int W, int C, int R, int stride,
                        int pad_t, int pad_l, int P, int Q, float* dw, void* ws, size_t ws_bytes,
                        void* stream);
/* squeeze-excite: out[b,c] = scale * sum_pix x[b,pix,c] * (y ? y[b,pix,c] : 1)   (adaptive_avg_pool2d
 * with scale = 1/HW, and the gate gradient); ws >= creste_chan_reduce_workspace_bytes(HW,C,1,B) */
int creste_sample_dot(const float* x, const float* y, int B, long long HW, int C, float scale, float* out,
                      void* ws, size_t ws_bytes, void* stream);
/* out[b,pix,c] = (x ? x*(a ? a[b,c] : 1) : 0) + (b ? b[b,c] : 0): the gate multiply and the pool adjoint */
int creste_sample_affine(const float* x, const float* a, const float* b, int B, long long HW, int C,
                         float* out, void* stream);
int creste_act(const float* x, long long n, int act, float* y, void* stream);
int creste_act_bwd(const float* g, const float* x, long long n, int act, float* dx, void* stream);
/* identity skip + drop-connect (efficientnet_pytorch drop_connect): out = inp + x*s[b] (s NULL: 1) */
int creste_add_scaled(const float* inp, const float* x, const float* s, int B, long long per, float* out,
                      void* stream);
/* out[pix,0:Cn] = x[pix,c0:c0+Cn]: adjoint of torch.cat([skip, up], 1) (effnet.py:24) */
int creste_chan_slice(const float* x, long long npix, int C, int c0, int Cn, float* out, void* stream);
/* weight gradient of the strided stem conv (C == 4): dw [R*S*C][K] */
size_t creste_wgrad_strided_workspace_bytes(int N, int P, int Q, int C, int K, int R, int S);
int creste_wgrad_strided(const float* x, const float* g, int N, int H, int W, int C, int K, int R, int S,
                         int stride, int pad_t, int pad_l, int P, int Q, float* dw, void* ws,
                         size_t ws_bytes, void* stream);
/* 1x1 weight gradient over a handful of rows (the squeeze-excite convs act on [B,1,1,C] vectors):
 * dw [C][K] = x^T g with x [npix,C], g [npix,K], npix <= 4096 */
int creste_wgrad_rows(const float* x, const float* g, int npix, int C, int K, float* dw, void* stream);
/* gradients of CrossEntropyDepth (loss_utils.py:477-527) w.r.t. the NCHW logits and of MSELoss
 * (:606-647); scale_dev is a DEVICE float (upstream gradient / #valid), so no host sync. */
int creste_ce_depth_bwd(const float* logits_nchw, const float* label_mm, int N, int D, long long HW,
                        float depth_min, float depth_max, const float* scale_dev, float* dlogits,
                        void* stream);
int creste_masked_mse_bwd(const float* pred, const float* gt, long long n, const float* scale_dev,
                          float* dpred, void* stream);

/* Weight gradient of a stride-1 dense conv on the tcgen05 tensor cores (3xFP16 split, fp32 accumulate):
 * same contract as creste_conv2d_wgrad (dw [R*S*C][K]) for the wide layers of the backbone
 * (C, K >= 64 and multiples of 8, N*P*Q >= 512).  GEMM over the PIXELS with MN-major operands staged by
 * TMA straight from the NHWC tensors (csrc/conv_tc.cu, wgrad_tc_kernel). */
int creste_conv2d_wgrad_tc_supported(const creste_conv_desc* d);
size_t creste_conv2d_wgrad_tc_workspace_bytes(const creste_conv_desc* d);
int creste_conv2d_wgrad_tc(const creste_conv_desc* d, const float* x, const float* g, float* dw, void* ws,
                           size_t ws_bytes, void* stream);
/* The 3xFP16 operand pre-pass of the tensor-core convs, stand-alone: amax -> power-of-two scale -> fp16 hi / lo of a
 * dense fp32 tensor (numel % 8 == 0); scal = DEVICE float[4] {s, 1/s, amax bits, -}.  Training: the forward conv's
 * operand is split ONCE and saved for the weight gradient, the output gradient is split ONCE for the data and the
 * weight gradient (creste_conv2d_presplit / creste_conv2d_wgrad_tc_presplit) -- 2 pre-passes per conv instead of 4. */
int creste_f16_split(const float* x, long long numel, void* hi, void* lo, float* scal, void* stream);
/* the same with max|x| known (DEVICE float[1], published by the kernel that produced x: creste_chan_affine_act_amax,
 * creste_chan_axpby_amax, the amax_out of creste_conv2d): one pass instead of three launches; bit-identical halves */
int creste_f16_split_amax(const float* x, long long numel, const float* amax, void* hi, void* lo, float* scal,
                          void* stream);
/* creste_conv2d_wgrad_tc on operands that are already split (same workspace size). */
int creste_conv2d_wgrad_tc_presplit(const creste_conv_desc* d, const void* x_hi, const void* x_lo, const float* x_scal,
                                    const void* g_hi, const void* g_lo, const float* g_scal, float* dw, void* ws,
                                    size_t ws_bytes, void* stream);

/* The 3xFP16 weight operand of creste_conv2d (precision 4) packed in ONE launch (training re-packs every step):
 * logical w[k][c][r][s] is read through element strides (sK, sC, sR, sS) -- the data-gradient conv passes the
 * transposed view of the flipped filter; out (floats) = npad*R*S*cpad64 fp16 hi halves, as many lo halves, then
 * npad fp32 inverse scales (npad from creste_conv2d_tc_layout, cpad64 = C rounded up to 64). */
int creste_pack_weight_f16(const float* w, long long sK, long long sC, long long sR, long long sS, int K, int C,
                           int R, int S, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stage-2 (train_ssc.py) losses: values and gradients.
 *
 * creste_smooth_l1: masked Smooth-L1 mean -- SmoothL1Depth on the soft-argmax depth (loss_utils.py:530-573:
 *   mask = valid LiDAR bins, gt_scale = 1/1000) and SmoothL1 on the elevation head (:576-604: mask NULL, non-finite
 *   targets skipped).  acc2 = DEVICE double {sum, count}; the backward writes scale * d smoothl1 / d pred.
 * creste_ce_weighted: class-weighted cross-entropy over masked cells (CrossEntropy, :379-474).  logits NCHW
 *   [B,C,HW], labels int64 [B,HW], mask uint8 [B,HW] (NULL = all), class_weights [C] (NULL = 1).
 *   acc4 = DEVICE double {sum w*nll, sum w, #argmax == label over label != 0, #(label != 0)}.
 * creste_supcon_*: multi-positive contrastive loss over L2-normalised embeddings (SupPixelConLoss :203-286 ->
 *   MultiPosConLoss, creste/models/losses/supcon_loss.py:56-115).  f [N,D] local rows, a [Na,D] all-gathered rows
 *   (row i of f is row i + self_off of a), labels int64; stats4 [N,4] float (row max, sum exp, #positives, sum of
 *   positive logits) is written by the forward and read by the backward; loss_sum = DEVICE double (sum over rows;
 *   the loss is loss_sum / N); df [N,D], da [Na,D] (the caller reduce-scatters da across ranks: the backward of
 *   torch.distributed.nn.all_gather).  D in {4, 8, 16, 32, 64, 128}.
 * creste_l2norm_rows: F.normalize(x, dim=-1) and its backward. */
int creste_smooth_l1(const float* pred, const float* gt, const uint8_t* mask, long long n, float gt_scale,
                     float beta, double* acc2, void* stream);
int creste_smooth_l1_bwd(const float* pred, const float* gt, const uint8_t* mask, long long n, float gt_scale,
                         float beta, const float* scale_dev, float* dpred, void* stream);
int creste_ce_weighted(const float* logits_nchw, const int64_t* labels, const uint8_t* mask,
                       const float* class_weights, int B, int C, long long HW, long long ignore_index, double* acc4,
                       void* stream);
int creste_ce_weighted_bwd(const float* logits_nchw, const int64_t* labels, const uint8_t* mask,
                           const float* class_weights, int B, int C, long long HW, long long ignore_index,
                           const float* scale_dev, float* dlogits, void* stream);
int creste_l2norm_rows(const float* x, int N, int D, float eps, float* y, float* norms, void* stream);
int creste_l2norm_rows_bwd(const float* y, const float* dy, const float* norms, int N, int D, float* dx,
                           void* stream);
int creste_supcon_fwd(const float* f, const float* a, const int64_t* lf, const int64_t* la, int N, int Na, int D,
                      int self_off, float temperature, const float* class_weights, float* stats4, double* loss_sum,
                      void* stream);
int creste_supcon_bwd(const float* f, const float* a, const int64_t* lf, const int64_t* la, int N, int Na, int D,
                      int self_off, float temperature, const float* class_weights, const float* stats4,
                      const float* scale_dev, float* df, float* da, void* stream);

/* ---- on-device input pipeline: the per-frame warps of the reference's CPU dataloader.
 * theta = DEVICE float[B][6]: the 2x3 map from normalised OUTPUT to normalised INPUT coordinates that kornia's
 * warp_affine hands to F.affine_grid (the host derives it from the pixel-space matrix: mirror
 * creste/utils/train_utils.py `affine_theta`).
 * creste_affine_warp: reference creste/utils/utils.py:6-38 `warp` (kornia warp_affine + the ones-channel validity
 *   mask thresholded at 0.99), used by RotateAndTranslate.transform_map (creste/utils/train_utils.py:213-232) and the
 *   FOV-mask pose warp (creste/datasets/codapefree_dataloader.py:691-709).  in [B][C][H][W] fp32 -> out
 *   [B][C][Ho][Wo], mask [B][Ho][Wo] uint8 (may be NULL); nearest: 0 = bilinear, 1 = nearest; zeros padding.
 * creste_depth_augment: DepthAugmentation.__call__ (creste/utils/train_utils.py:110-181) in one pass given its draws:
 *   out = warp_bilinear(depth * (u > p_drop), theta, align_corners = true) + g * noise_std, all [H][W] fp32.
 * creste_traverse_to_bev: CodaPEFreeDataset._load_traverse (creste/datasets/codapefree_dataloader.py:590-615):
 *   poses [T][4][4] (LiDAR frame, relative to the first) -> clamped BEV grid poses [T][3][3]. */
int creste_affine_warp(const float* in, int B, int C, int H, int W, const float* theta, int Ho, int Wo, int nearest,
                       int align_corners, float* out, unsigned char* mask, void* stream);
int creste_depth_augment(const float* depth, const float* u, const float* g, int H, int W, float p_drop,
                         const float* theta, float noise_std, float* out, void* stream);
int creste_traverse_to_bev(const float* poses, int T, float voxel_x, float voxel_y, int bev_h, int bev_w, float* out,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRESTE_B200_H */
