"""Tensor-level wrappers over the C ABI (include/creste_b200.h).

PyTorch is used for device memory and streams only: every function here takes CUDA tensors,
allocates outputs / scratch with torch's caching allocator on the same device and launches the
sm_100a kernels on torch's current stream.  No function has a CPU or PyTorch-eager fallback.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import ConvDesc, check, lib, ptr, stream

ACT = {"none": 0, None: 0, "relu": 1, "swish": 2, "sigmoid": 3}
PRECISION = {"fp32": 0, "3xtf32": 1, "tf32": 2, "bf16": 3, "3xfp16": 4, "fp16": 5}


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------ VI
def vi_solve(r, gamma=0.99, thr=1e-3, max_sweeps=4096, want_q=True):
    """Value iteration (reference vin.py:48-80).  r [B,1,H,W] or [B,H,W] fp32 CUDA.
    Returns v [B,1,H,W], q [B,8,H,W], pi [B,8,H,W], info int32[2] (device: sweeps, hit_max)."""
    r3 = r.reshape(r.shape[0], r.shape[-2], r.shape[-1]).contiguous().float()
    B, H, W = r3.shape
    v = torch.empty_like(r3)
    q = torch.empty(B, 8, H, W, device=r.device) if want_q else None
    pi = torch.empty(B, 8, H, W, device=r.device) if want_q else None
    info = torch.zeros(2, dtype=torch.int32, device=r.device)
    n = lib().creste_vi_workspace_bytes(B, H, W, max_sweeps)
    ws = _ws(n, r.device)
    check(lib().creste_vi_solve(ptr(r3), ptr(v), ptr(q), ptr(pi), B, H, W, C.c_float(gamma),
                                C.c_float(thr), max_sweeps, ptr(info), ptr(ws), C.c_size_t(n),
                                stream()), "creste_vi_solve")
    return v.view(B, 1, H, W), q, pi, info


# ----------------------------------------------------------------------------------------- SVF
def svf(policy, expert_rc, fov, T, ds=2, sharpen=True, temperature=0.005, zero_terminal=False):
    """Expected SVF + greedy rollout (reference lfd.py:156-277).  policy [B,8,H,W];
    expert_rc [B,Te,2] fp32 (Te expert poses; T = action horizon); fov [H,W] uint8/bool.
    Returns exp_svf [B,H,W], states [B,T,2] int64, states_grid [B,H,W]."""
    policy = policy.contiguous().float()
    expert_rc = expert_rc.contiguous().float()
    fov = fov.to(torch.uint8).contiguous()
    B, A, H, W = policy.shape
    Te = expert_rc.shape[1]
    assert A == 8 and tuple(expert_rc.shape) == (B, Te, 2) and tuple(fov.shape) == (H, W)
    out = torch.empty(B, H, W, device=policy.device)
    states = torch.empty(B, T, 2, dtype=torch.int64, device=policy.device)
    grid = torch.empty(B, H, W, device=policy.device)
    n = lib().creste_svf_workspace_bytes(B, H, W, T)
    ws = _ws(n, policy.device)
    check(lib().creste_svf(ptr(policy), ptr(expert_rc), ptr(fov), B, H, W, T, Te, ds,
                           int(bool(sharpen)), C.c_float(temperature), int(bool(zero_terminal)),
                           ptr(out), ptr(states), ptr(grid), ptr(ws), C.c_size_t(n), stream()),
          "creste_svf")
    return out, states, grid


# --------------------------------------------------------------------------------------- splat
def frustum_to_bev(depth, p2p, pc_range, voxel):
    """depth [N,Hs,Ws] m, p2p [N,4,4] -> xy [N,P,2], z [N,P], mask [N,P] uint8
    (reference splat_projection.py:19-51, :169, :175-189).  pc_range/voxel: python floats."""
    depth = depth.contiguous().float()
    p2p = p2p.contiguous().float()
    N, Hs, Ws = depth.shape
    P = Hs * Ws
    xy = torch.empty(N, P, 2, device=depth.device)
    z = torch.empty(N, P, device=depth.device)
    mask = torch.empty(N, P, dtype=torch.uint8, device=depth.device)
    rng = (C.c_float * 6)(*[float(v) for v in pc_range])
    vox = (C.c_float * 2)(float(voxel[0]), float(voxel[1]))
    check(lib().creste_frustum_to_bev(ptr(depth), ptr(p2p), N, Hs, Ws, rng, vox, ptr(xy), ptr(z),
                                      ptr(mask), stream()), "creste_frustum_to_bev")
    return xy, z, mask


def camera_to_world(depth, p2p):
    """depth [N,Hs,Ws], p2p [N,4,4] -> xyz [N,3,Hs,Ws] (reference splat_projection.py:19-51)."""
    depth = depth.contiguous().float()
    p2p = p2p.contiguous().float()
    N, Hs, Ws = depth.shape
    xyz = torch.empty(N, 3, Hs, Ws, device=depth.device)
    check(lib().creste_camera_to_world(ptr(depth), ptr(p2p), N, Hs, Ws, ptr(xyz), stream()),
          "creste_camera_to_world")
    return xyz


def points_to_voxels(points, lidar2map, voxel):
    """points [B,P,3] CUDA; lidar2map 4x4 / voxel (x, y) python floats -> voxels [B,P,2]
    (reference splat_projection.py:175-189)."""
    points = points.contiguous().float()
    B, P, _ = points.shape
    xy = torch.empty(B, P, 2, device=points.device)
    L = (C.c_float * 16)(*[float(v) for row in lidar2map for v in row])
    vox = (C.c_float * 2)(float(voxel[0]), float(voxel[1]))
    check(lib().creste_points_to_voxels(ptr(points), C.c_longlong(B * P), L, vox, ptr(xy), stream()),
          "creste_points_to_voxels")
    return xy


def zmlp_concat(feats_nhwc, z, w1, b1, w2, b2, amax_out=None):
    """feats NHWC [...,C] + MLP(z) -> NHWC [...,C+32] (reference splat_projection.py:152-158).
    amax_out: optional device float[1] that receives max|out|."""
    Cc = feats_nhwc.shape[-1]
    NP = feats_nhwc.numel() // Cc
    out = torch.empty(*feats_nhwc.shape[:-1], Cc + 32, device=feats_nhwc.device)
    check(lib().creste_zmlp_concat_ex(ptr(feats_nhwc), ptr(z.contiguous()), NP, Cc, ptr(w1), ptr(b1),
                                      ptr(w2), ptr(b2), ptr(out), ptr(amax_out), stream()), "creste_zmlp_concat")
    return out


def splat_soft(xy, feats_nhwc, mask, H, W, min_weight=1.0, want_nhwc=True, want_nchw=True,
               want_idx=False):
    """Bilinear splat (reference splat_projection.py:262-354).  xy [N,P,2]; feats NHWC [N,P,F];
    mask [N,P] uint8 or None.  Returns dict(bev_nhwc [N,H,W,F], bev_nchw [N,F,H,W],
    dens [N,1,H,W], idx [N,P,4] int64)."""
    xy = xy.contiguous().float()
    N, P, _ = xy.shape
    dev = xy.device
    F = 0
    if feats_nhwc is not None:
        feats_nhwc = feats_nhwc.contiguous().float()
        F = feats_nhwc.shape[-1]
        assert feats_nhwc.numel() == N * P * F
    want_feats = feats_nhwc is not None and (want_nhwc or want_nchw)
    nhwc = torch.empty(N, H, W, F, device=dev) if (want_feats and want_nhwc) else None
    nchw = torch.empty(N, F, H, W, device=dev) if (want_feats and want_nchw) else None
    dens = torch.empty(N, 1, H, W, device=dev)
    idx = torch.empty(N, P, 4, dtype=torch.int64, device=dev) if want_idx else None
    n = lib().creste_splat_workspace_bytes(N, H, W, max(F, 1)) if want_feats else 0
    ws = _ws(n, dev)
    check(lib().creste_splat_soft(ptr(xy), ptr(feats_nhwc), ptr(mask), N, P, F, H, W,
                                  C.c_float(min_weight), ptr(nhwc), ptr(nchw), ptr(dens), ptr(idx),
                                  ptr(ws), C.c_size_t(n), stream()), "creste_splat_soft")
    return {"bev_nhwc": nhwc, "bev_nchw": nchw, "dens": dens, "idx": idx}


def splat_soft_bwd(xy, feats_nhwc, mask, bev_nhwc, dens, g_bev_nhwc, g_dens, min_weight=1.0):
    """Backward of splat_soft: -> (dfeats [N,P,F], dxy [N,P,2]).  bev_nhwc [N,H,W,F] / dens [N,1,H,W] are the
    forward outputs, g_bev_nhwc / g_dens (None = 0) their gradients (reference autograd of :262-354)."""
    xy, feats_nhwc = xy.contiguous().float(), feats_nhwc.contiguous().float()
    N, P, _ = xy.shape
    F = feats_nhwc.shape[-1]
    _, H, W, _ = bev_nhwc.shape
    dfeats = torch.empty(N, P, F, device=xy.device)
    dxy = torch.empty(N, P, 2, device=xy.device)
    n = lib().creste_splat_bwd_workspace_bytes(N, H, W)
    ws = _ws(n, xy.device)
    check(lib().creste_splat_soft_bwd(ptr(xy), ptr(feats_nhwc), ptr(mask), ptr(bev_nhwc.contiguous()),
                                      ptr(dens.contiguous()), ptr(g_bev_nhwc.contiguous().float()),
                                      ptr(None if g_dens is None else g_dens.contiguous().float()), N, P, F, H, W,
                                      C.c_float(min_weight), ptr(dfeats), ptr(dxy), ptr(ws), C.c_size_t(n), stream()),
          "creste_splat_soft_bwd")
    return dfeats, dxy


def frustum_bwd(dxy, dz, p2p, shape, voxel):
    """-> d depth [N,Hs,Ws] from d xy [N,P,2] and / or d z [N,P] (either may be None)."""
    N, Hs, Ws = shape
    p2p = p2p.contiguous().float()
    out = torch.empty(N, Hs, Ws, device=p2p.device)
    vox = (C.c_float * 2)(float(voxel[0]), float(voxel[1]))
    check(lib().creste_frustum_bwd(ptr(None if dxy is None else dxy.contiguous().float()),
                                   ptr(None if dz is None else dz.contiguous().float()), ptr(p2p), N, Hs, Ws, vox,
                                   ptr(out), stream()), "creste_frustum_bwd")
    return out


def depth_expectation_bwd(logits_nhwc, g_metric, dmin=300.0, dmax=25600.0, out_div=1000.0):
    logits_nhwc = logits_nhwc.contiguous()
    D = logits_nhwc.shape[-1]
    NP = logits_nhwc.numel() // D
    out = torch.empty_like(logits_nhwc)
    check(lib().creste_depth_expectation_bwd(ptr(logits_nhwc), ptr(g_metric.contiguous().float()), NP, D,
                                             C.c_float(dmin), C.c_float(dmax), C.c_float(out_div), ptr(out), stream()),
          "creste_depth_expectation_bwd")
    return out


def dilate(g_nhwc, stride, Hz, Wz):
    """Zero insertion: z[n, p*stride, q*stride, :] = g[n,p,q,:] on an [N,Hz,Wz,C] zero canvas."""
    g_nhwc = g_nhwc.contiguous()
    N, P, Q, Cc = g_nhwc.shape
    z = torch.empty(N, Hz, Wz, Cc, device=g_nhwc.device)
    check(lib().creste_dilate(ptr(g_nhwc), N, P, Q, Cc, int(stride), int(Hz), int(Wz), ptr(z), stream()), "creste_dilate")
    return z


def phase_slice(x_nhwc, stride, a, b, Ha, Wa):
    """x[:, a::stride, b::stride, :] padded / cropped to [N,Ha,Wa,C] (zeros outside the image)."""
    x_nhwc = x_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    out = torch.empty(N, Ha, Wa, Cc, device=x_nhwc.device)
    check(lib().creste_phase_slice(ptr(x_nhwc), N, H, W, Cc, int(stride), int(a), int(b), int(Ha), int(Wa), ptr(out),
                                   stream()), "creste_phase_slice")
    return out


# --------------------------------------------------------------------------------------- LiDAR
def lidar_raster(pc, P34, H, W, out_mm=None, want_m=True):
    """pc [n,>=3] fp32 CUDA; P34 [3,4] float64 (host array-like) -> depth_m [H,W], depth_mm [H,W]
    (reference projection.py:64-134 + build_dense_depth.py:461-463).  `out_mm`: optional
    contiguous [H,W] view to write the millimetre raster into (e.g. channel 3 of the RGB-D
    network input)."""
    pc = pc.contiguous().float()
    flat = [float(v) for row in P34 for v in row]
    assert len(flat) == 12
    Pm = (C.c_double * 12)(*flat)
    dm = torch.empty(H, W, device=pc.device) if want_m else None
    dmm = out_mm if out_mm is not None else torch.empty(H, W, device=pc.device)
    assert tuple(dmm.shape) == (H, W)
    ws = _ws(H * W * 8, pc.device)
    check(lib().creste_lidar_raster(ptr(pc), pc.shape[0], pc.shape[1], Pm, H, W, ptr(dm), ptr(dmm),
                                    ptr(ws), C.c_size_t(H * W * 8), stream()), "creste_lidar_raster")
    return dm, dmm


def depth_expectation(logits_nhwc, dmin=300.0, dmax=25600.0, out_div=1000.0):
    """logits NHWC [N,Hs,Ws,128] -> metric [N,Hs,Ws] (expectation / out_div: metres by default),
    bins [N,Hs,Ws] int64 (reference depth_utils.py:300-313, depth.py:60-100)."""
    D = logits_nhwc.shape[-1]
    shp = logits_nhwc.shape[:-1]
    NP = logits_nhwc.numel() // D
    metric = torch.empty(shp, device=logits_nhwc.device)
    bins = torch.empty(shp, dtype=torch.int64, device=logits_nhwc.device)
    check(lib().creste_depth_expectation(ptr(logits_nhwc), NP, D, C.c_float(dmin), C.c_float(dmax),
                                         C.c_float(out_div), ptr(metric), ptr(bins), stream()),
          "creste_depth_expectation")
    return metric, bins


BIN_MODES = {"UD": 0, "LID": 1, "SID": 2}


def bin_depths(depth, mode, depth_min, depth_max, num_bins, target=False):
    """Depth map -> (float | int64) bin indices, reference depth_utils.py:346-383."""
    d = depth.contiguous().float()
    out = torch.empty(d.shape, dtype=torch.int64 if target else torch.float32, device=d.device)
    check(lib().creste_bin_depths(ptr(d), C.c_longlong(d.numel()), BIN_MODES[mode], C.c_float(depth_min),
                                  C.c_float(depth_max), int(num_bins), int(bool(target)),
                                  ptr(None if target else out), ptr(out if target else None), stream()),
          "creste_bin_depths")
    return out


# ---------------------------------------------------------------------------------- conv family
def pack_conv_weight(w):
    """[K,C,R,S] (torch layout) -> [R*S*C, ldw] row-major, ldw = K rounded up to 4."""
    K, Cc, R, S = w.shape
    ldw = (K + 3) // 4 * 4
    p = w.permute(2, 3, 1, 0).reshape(R * S * Cc, K)
    if ldw != K:
        p = torch.cat([p, p.new_zeros(R * S * Cc, ldw - K)], dim=1)
    return p.contiguous()


def rna_tf32(t):
    """Round fp32 to tf32 (10-bit mantissa), nearest with ties away from zero = cvt.rna.tf32.f32."""
    u = t.contiguous().view(torch.int32)
    return ((u + 0x1000) & ~0x1FFF).view(torch.float32)


def tc_layout(K, Cc, R, S):
    bn, npad, cpad = C.c_int(0), C.c_int(0), C.c_int(0)
    lib().creste_conv2d_tc_layout(K, Cc, R, S, C.byref(bn), C.byref(npad), C.byref(cpad))
    return bn.value, npad.value, cpad.value


def pack_conv_weight_tc(w, split):
    """[K,C,R,S] -> [Npad][R*S*Cpad] tf32-rounded hi (+ lo when split), one flat buffer."""
    K, Cc, R, S = w.shape
    _, npad, cpad = tc_layout(K, Cc, R, S)
    p = w.new_zeros(npad, R * S, cpad)
    p[:K, :, :Cc] = w.permute(0, 2, 3, 1).reshape(K, R * S, Cc)
    hi = rna_tf32(p)
    if not split:
        return hi.reshape(-1).contiguous()
    lo = rna_tf32(p - hi)
    return torch.cat([hi.reshape(-1), lo.reshape(-1)]).contiguous()


def pack_conv_weight_f16(w):
    """[K,C,R,S] -> the 3xFP16 operand of creste_conv2d (precision 4): per-output-channel
    power-of-two scale s_k with max_c|w_k| * s_k in [2^14, 2^15); hi = fp16(w * s_k),
    lo = fp16((w * s_k - hi) * 2^11); layout [Npad][R*S*Cpad64] hi halves, then lo halves, then
    1/s_k as fp32 [Npad] -- returned as one flat float32 buffer."""
    K, Cc, R, S = w.shape
    _, npad, _ = tc_layout(K, Cc, R, S)
    cpad = (Cc + 63) // 64 * 64
    p = w.new_zeros(npad, R * S, cpad)
    p[:K, :, :Cc] = w.permute(0, 2, 3, 1).reshape(K, R * S, Cc)
    amax = p.abs().amax(dim=(1, 2))
    e = torch.floor(torch.log2(torch.where(amax > 0, amax, torch.ones_like(amax))))
    k = (14 - e).clamp(-100, 100)
    sc = torch.exp2(k).view(npad, 1, 1)
    ps = p * sc
    hi = ps.half()
    lo = ((ps - hi.float()) * 2048.0).half()
    halves = torch.cat([hi.reshape(-1), lo.reshape(-1)])
    return torch.cat([halves.view(torch.float32), torch.exp2(-k).float()]).contiguous()


def conv_desc(x_shape, K, R, S, stride, pad, act="none", out_nchw=False, precision="fp32"):
    N, H, W, Cc = x_shape
    pt, pb, pl, pr = pad
    P = (H + pt + pb - R) // stride + 1
    Q = (W + pl + pr - S) // stride + 1
    return ConvDesc(N, H, W, Cc, K, R, S, stride, pt, pl, P, Q, ACT[act], int(out_nchw),
                    PRECISION[precision])


def tc_supported(x_shape, K, R, S, stride, pad, precision):
    if PRECISION[precision] == 0:
        return False
    d = conv_desc(x_shape, K, R, S, stride, pad, precision=precision)
    return bool(lib().creste_conv2d_tc_supported(C.byref(d)))


def conv2d(x_nhwc, w_packed, K, R, S, stride=1, pad=(0, 0, 0, 0), scale=None, shift=None,
           gate=None, residual=None, act="none", out_nchw=False, precision="fp32", amax_in=None, amax_out=None,
           split_out=None):
    """NHWC conv + folded BN/bias + residual + activation.  pad = (top, bottom, left, right).
    amax_in / amax_out: optional device float[1] tensors -- an upper bound of max|x * gate| that lets the 3xFP16
    operand pre-pass skip its amax pass, and the slot that receives max|out| from the epilogue.
    split_out = (mode, bound_mul, bound_add), mode "only" | "both": the output is (also) written as the next
    tensor-core conv's SplitAct operand from the epilogue; returns the SplitAct ("only") or (out, SplitAct)."""
    N, H, W, Cc = x_nhwc.shape
    pt, pb, pl, pr = pad
    P = (H + pt + pb - R) // stride + 1
    Q = (W + pl + pr - S) // stride + 1
    d = ConvDesc(N, H, W, Cc, K, R, S, stride, pt, pl, P, Q, ACT[act], int(out_nchw),
                 PRECISION[precision])
    n = lib().creste_conv2d_workspace_bytes(C.byref(d))
    ws = _ws(n, x_nhwc.device) if n else None
    if split_out is not None:
        mode, bmul, badd = split_out
        so = _new_split((N, P, Q, K), x_nhwc.device, want_lo=(precision == "3xfp16"))
        out = torch.empty(N, P, Q, K, device=x_nhwc.device) if mode == "both" else None
        check(lib().creste_conv2d_split_out(C.byref(d), ptr(x_nhwc), ptr(w_packed), ptr(scale), ptr(shift), ptr(gate),
                                            ptr(residual), ptr(out), ptr(amax_in), ptr(amax_out), ptr(so.hi), ptr(so.lo),
                                            ptr(so.scal), C.c_float(bmul), C.c_float(badd), ptr(ws), C.c_size_t(n),
                                            stream()), "creste_conv2d_split_out")
        return so if out is None else (out, so)
    out = torch.empty((N, K, P, Q) if out_nchw else (N, P, Q, K), device=x_nhwc.device)
    check(lib().creste_conv2d_ex(C.byref(d), ptr(x_nhwc), ptr(w_packed), ptr(scale), ptr(shift),
                                 ptr(gate), ptr(residual), ptr(out), ptr(amax_in), ptr(amax_out), ptr(ws),
                                 C.c_size_t(n), stream()), "creste_conv2d")
    return out


def _new_split(shape, device, want_lo=True):
    hi = torch.empty(shape, dtype=torch.float16, device=device)
    return SplitAct(hi, torch.empty_like(hi) if want_lo else None, torch.empty(2, device=device), shape)


class SplitAct:
    """An activation tensor held as the 3xFP16 operand of a tensor-core conv: fp16 hi / lo halves [N,H,W,C] and the
    device scale record {s, 1/s}.  Produced by upsample_concat_split, consumed by conv2d_presplit."""

    def __init__(self, hi, lo, scal, shape, amax=None):
        self.hi, self.lo, self.scal, self.shape = hi, lo, scal, tuple(shape)
        self.device = hi.device
        self.amax = amax          # device float[1]: the true max|x| of the fp32 values, when the producer measured it


def upsample_concat_split(skip_nhwc, x_nhwc, out_hw, scale_factor, amax_skip, amax_x, x_first=False, want_lo=True):
    """upsample_concat whose output is written directly as a SplitAct (no fp32 tensor, no split pre-pass)."""
    N, Hi, Wi, Cx = x_nhwc.shape
    Ho, Wo = out_hw
    if scale_factor is None:
        rh, rw = Hi / Ho, Wi / Wo
    else:
        sh, sw = (scale_factor, scale_factor) if not isinstance(scale_factor, (tuple, list)) else scale_factor
        rh, rw = 1.0 / sh, 1.0 / sw
    Cs = 0 if skip_nhwc is None else skip_nhwc.shape[-1]
    dev = x_nhwc.device
    hi = torch.empty(N, Ho, Wo, Cs + Cx, dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi) if want_lo else None
    scal = torch.empty(2, device=dev)
    a, b = (amax_x, None) if skip_nhwc is None else (amax_skip, amax_x)
    check(lib().creste_upsample_concat_split(ptr(skip_nhwc), Cs, ptr(x_nhwc), N, Hi, Wi, Cx, Ho, Wo, C.c_float(rh),
                                             C.c_float(rw), int(bool(x_first)), ptr(a), ptr(b), ptr(hi), ptr(lo),
                                             ptr(scal), stream()), "creste_upsample_concat_split")
    return SplitAct(hi, lo, scal, (N, Ho, Wo, Cs + Cx))


def conv2d_presplit(xs, w_packed, K, R, S, stride=1, pad=(0, 0, 0, 0), scale=None, shift=None, residual=None,
                    act="none", out_nchw=False, precision="3xfp16", amax_out=None, split_out=None):
    """Tensor-core conv on a SplitAct input (no activation pre-pass).  split_out: as in conv2d."""
    N, H, W, Cc = xs.shape
    pt, pb, pl, pr = pad
    P = (H + pt + pb - R) // stride + 1
    Q = (W + pl + pr - S) // stride + 1
    d = ConvDesc(N, H, W, Cc, K, R, S, stride, pt, pl, P, Q, ACT[act], int(out_nchw), PRECISION[precision])
    if split_out is not None:
        mode, bmul, badd = split_out
        so = _new_split((N, P, Q, K), xs.device, want_lo=(precision == "3xfp16"))
        out = torch.empty(N, P, Q, K, device=xs.device) if mode == "both" else None
        check(lib().creste_conv2d_presplit_split_out(C.byref(d), ptr(xs.hi), ptr(xs.lo), ptr(xs.scal), ptr(w_packed),
                                                     ptr(scale), ptr(shift), ptr(residual), ptr(out), ptr(amax_out),
                                                     ptr(so.hi), ptr(so.lo), ptr(so.scal), C.c_float(bmul),
                                                     C.c_float(badd), ptr(xs.amax), stream()),
              "creste_conv2d_presplit_split_out")
        return so if out is None else (out, so)
    out = torch.empty((N, K, P, Q) if out_nchw else (N, P, Q, K), device=xs.device)
    check(lib().creste_conv2d_presplit(C.byref(d), ptr(xs.hi), ptr(xs.lo), ptr(xs.scal), ptr(w_packed), ptr(scale),
                                       ptr(shift), ptr(residual), ptr(out), ptr(amax_out), stream()),
          "creste_conv2d_presplit")
    return out


def dwconv_bn_swish(x_nhwc, w_rsc, scale, shift, R, stride, pad, amax_out=None):
    """Depthwise conv + BN + swish; returns (out NHWC, chan_part [N,nparts,C] SE partial sums).
    amax_out: optional device float[1] that receives max|out|."""
    N, H, W, Cc = x_nhwc.shape
    pt, pb, pl, pr = pad
    P = (H + pt + pb - R) // stride + 1
    Q = (W + pl + pr - R) // stride + 1
    out = torch.empty(N, P, Q, Cc, device=x_nhwc.device)
    # the chan_part row count selects the kernel: tiled (shared-memory input tile) or x-blocked, whichever is faster
    nparts = lib().creste_dwconv_parts(N, P, Q, R, stride)
    csum = torch.empty(N, nparts, Cc, device=x_nhwc.device)
    check(lib().creste_dwconv_bn_swish_ex(ptr(x_nhwc), ptr(w_rsc), ptr(scale), ptr(shift), N, H, W, Cc,
                                          R, stride, pt, pl, P, Q, ptr(out), ptr(csum), nparts, ptr(amax_out), stream()),
          "creste_dwconv_bn_swish")
    return out, csum


def se_gate(chan_part, hw, w_red, b_red, w_exp, b_exp):
    N, nparts, Cc = chan_part.shape
    Csq = w_red.shape[0]
    gate = torch.empty(N, Cc, device=chan_part.device)
    check(lib().creste_se_gate(ptr(chan_part), nparts, C.c_float(1.0 / hw), N, Cc, Csq, ptr(w_red), ptr(b_red),
                               ptr(w_exp), ptr(b_exp), ptr(gate), stream()), "creste_se_gate")
    return gate


def upsample_concat(skip_nhwc, x_nhwc, out_hw, scale_factor=None, x_first=False):
    """cat([skip, bilinear(x)], C) in NHWC.  scale_factor: the nn.Upsample argument (number or
    (sh, sw)); when given, the sampling ratio is 1/scale_factor exactly as PyTorch does."""
    N, Hi, Wi, Cx = x_nhwc.shape
    Ho, Wo = out_hw
    if scale_factor is None:
        rh, rw = Hi / Ho, Wi / Wo
    else:
        sh, sw = (scale_factor, scale_factor) if not isinstance(scale_factor, (tuple, list)) \
            else scale_factor
        rh, rw = 1.0 / sh, 1.0 / sw
    Cs = 0 if skip_nhwc is None else skip_nhwc.shape[-1]
    out = torch.empty(N, Ho, Wo, Cs + Cx, device=x_nhwc.device)
    check(lib().creste_upsample_concat(ptr(skip_nhwc), Cs, ptr(x_nhwc), N, Hi, Wi, Cx, Ho, Wo,
                                       C.c_float(rh), C.c_float(rw), int(bool(x_first)), ptr(out),
                                       stream()),
          "creste_upsample_concat")
    return out


def maxpool2_concat(srcs_nhwc, rows_out=None, want_nchw=False):
    """2x2/2 max-pool of the channel concat of NHWC sources, cropped to `rows_out` output rows."""
    N, H, W, _ = srcs_nhwc[0].shape
    chans = [int(s.shape[-1]) for s in srcs_nhwc]
    rows_out = H // 2 if rows_out is None else rows_out
    Ct = sum(chans)
    out = torch.empty(N, rows_out, W // 2, Ct, device=srcs_nhwc[0].device)
    nchw = torch.empty(N, Ct, rows_out, W // 2, device=out.device) if want_nchw else None
    ps = (C.c_void_p * len(srcs_nhwc))(*[ptr(s).value for s in srcs_nhwc])
    cs = (C.c_int * len(chans))(*chans)
    check(lib().creste_maxpool2_concat(ps, cs, len(chans), N, H, W, rows_out, ptr(out), ptr(nchw),
                                       stream()), "creste_maxpool2_concat")
    return (out, nchw) if want_nchw else out


def nchw_to_nhwc(x):
    N, Cc, H, W = x.shape
    if Cc == 1 or H * W == 1:        # same memory order either way: a copy, not a transpose (the reward map has C = 1;
        return x.contiguous().view(N, H, W, Cc).clone()      # the tiled kernel ran 16 K one-row tiles for it: 190 us)
    out = torch.empty(N, H, W, Cc, device=x.device)
    check(lib().creste_nchw_to_nhwc(ptr(x.contiguous()), N, Cc, H, W, ptr(out), stream()),
          "creste_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x):
    N, H, W, Cc = x.shape
    if Cc == 1 or H * W == 1:
        return x.contiguous().view(N, Cc, H, W).clone()
    out = torch.empty(N, Cc, H, W, device=x.device)
    check(lib().creste_nhwc_to_nchw(ptr(x.contiguous()), N, H, W, Cc, ptr(out), stream()),
          "creste_nhwc_to_nchw")
    return out


def affine_warp(x, theta, out_hw=None, nearest=False, align_corners=False, want_mask=True):
    """kornia warp_affine / the reference's `warp` on the device: x [B,C,H,W] fp32, theta [B,2,3] (normalised output
    -> normalised input coordinates) -> (out [B,C,Ho,Wo], mask [B,Ho,Wo] bool | None)."""
    B, Cc, H, W = x.shape
    Ho, Wo = (H, W) if out_hw is None else (int(out_hw[0]), int(out_hw[1]))
    x = x.float().contiguous()
    theta = theta.to(device=x.device, dtype=torch.float32).reshape(B, 6).contiguous()
    out = torch.empty(B, Cc, Ho, Wo, device=x.device)
    mask = torch.empty(B, Ho, Wo, dtype=torch.uint8, device=x.device) if want_mask else None
    check(lib().creste_affine_warp(ptr(x), B, Cc, H, W, ptr(theta), Ho, Wo, int(bool(nearest)), int(bool(align_corners)),
                                   ptr(out), ptr(mask), stream()), "creste_affine_warp")
    return out, (mask.bool() if want_mask else None)


def depth_augment(depth, u, g, theta, dropout_prob, noise_std):
    """DepthAugmentation.__call__ given its draws, one pass: depth / u / g [1,H,W] or [H,W] fp32, theta [2,3]."""
    shp = depth.shape
    H, W = int(shp[-2]), int(shp[-1])
    depth, u, g = (t.float().contiguous() for t in (depth, u, g))
    theta = theta.to(device=depth.device, dtype=torch.float32).reshape(6).contiguous()
    out = torch.empty(shp, device=depth.device)
    check(lib().creste_depth_augment(ptr(depth), ptr(u), ptr(g), H, W, C.c_float(dropout_prob), ptr(theta),
                                     C.c_float(noise_std), ptr(out), stream()), "creste_depth_augment")
    return out


def traverse_to_bev(lidar_poses, voxel_size, bev_size):
    """CodaPEFreeDataset._load_traverse from the relative LiDAR poses [T,4,4] -> clamped BEV grid poses [T,3,3]."""
    P = lidar_poses.float().contiguous()
    T = P.shape[0]
    out = torch.empty(T, 3, 3, device=P.device)
    check(lib().creste_traverse_to_bev(ptr(P), T, C.c_float(float(voxel_size[0])), C.c_float(float(voxel_size[1])),
                                       int(bev_size[0]), int(bev_size[1]), ptr(out), stream()), "creste_traverse_to_bev")
    return out


def proj_head(x_nhwc, w_kc, bias, want_nchw=True, want_x_nchw=True):
    """1x1 conv C -> K (K <= 32) + the NCHW copies the output dict wants, one pass over x.
    -> (pred NHWC [N,H,W,K], pred NCHW [N,K,H,W] | None, x NCHW [N,C,H,W] | None)."""
    x_nhwc = x_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    K = w_kc.shape[0]
    dev = x_nhwc.device
    pred = torch.empty(N, H, W, K, device=dev)
    pred_nchw = torch.empty(N, K, H, W, device=dev) if want_nchw else None
    x_nchw = torch.empty(N, Cc, H, W, device=dev) if want_x_nchw else None
    check(lib().creste_proj_head(ptr(x_nhwc), ptr(w_kc.contiguous()), ptr(bias), N, H, W, Cc, K, ptr(pred),
                                 ptr(pred_nchw), ptr(x_nchw), stream()), "creste_proj_head")
    return pred, pred_nchw, x_nchw


def expert_visitation(traj_rc, map_ds, max_steps, H, W):
    """traj_rc [B,T,2] fp32 or fp64 CUDA -> counts [B,H,W] (reference loss_utils.py:1055-1116)."""
    traj_rc = traj_rc.contiguous()
    assert traj_rc.dtype in (torch.float32, torch.float64)
    B, T, _ = traj_rc.shape
    counts = torch.empty(B, H, W, device=traj_rc.device)
    check(lib().creste_expert_visitation(ptr(traj_rc), int(traj_rc.dtype == torch.float64), B, T,
                                         C.c_double(float(map_ds)), int(max_steps), H, W,
                                         ptr(counts), stream()), "creste_expert_visitation")
    return counts


# ------------------------------------------------------------------ stage-3 training primitives
def chan_affine(x, a=None, b=None, relu=False):
    """y[..., c] = act(x[..., c] * a[c] + b[c]) over a channels-last tensor (a/b None = 1/0)."""
    x = x.contiguous()
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    if _want_amax(x):
        amax = torch.empty(1, device=x.device)
        check(lib().creste_chan_affine_amax(ptr(x), ptr(a), ptr(b), C.c_longlong(x.numel() // Cc), Cc,
                                            int(bool(relu)), ptr(y), ptr(amax), stream()), "creste_chan_affine_amax")
        return publish_amax(y, amax)
    check(lib().creste_chan_affine(ptr(x), ptr(a), ptr(b), C.c_longlong(x.numel() // Cc), Cc,
                                   int(bool(relu)), ptr(y), stream()), "creste_chan_affine")
    return y


def relu_bwd(g, y):
    g, y = g.contiguous(), y.contiguous()
    out = torch.empty_like(g)
    if _want_amax(g):
        amax = torch.empty(1, device=g.device)
        check(lib().creste_relu_bwd_amax(ptr(g), ptr(y), C.c_longlong(g.numel()), ptr(out), ptr(amax), stream()),
              "creste_relu_bwd_amax")
        return publish_amax(out, amax)
    check(lib().creste_relu_bwd(ptr(g), ptr(y), C.c_longlong(g.numel()), ptr(out), stream()),
          "creste_relu_bwd")
    return out


def chan_dot(x, y=None):
    """out[c] = sum over all leading dims of x[..., c] * y[..., c]  (y None: plain channel sum)."""
    x = x.contiguous()
    y = None if y is None else y.contiguous()
    Cc = x.shape[-1]
    npix = x.numel() // Cc
    if Cc > 1024 and Cc % 4 == 0:       # the widest EfficientNet mid tensors (1152 channels)
        return sample_dot(x.view(1, npix, Cc), None if y is None else y.view(1, npix, Cc)).view(Cc)
    out = torch.empty(Cc, device=x.device)
    lib().creste_chan_dot_workspace_bytes.restype = C.c_size_t
    n = lib().creste_chan_dot_workspace_bytes(C.c_longlong(npix), Cc)
    ws = _ws(n, x.device)
    check(lib().creste_chan_dot(ptr(x), ptr(y), C.c_longlong(npix), Cc, ptr(out), ptr(ws),
                                C.c_size_t(n), stream()), "creste_chan_dot")
    return out


def chan_stats(x):
    """-> float64 [2, C]: per-channel sum and sum of squares over all leading dims (one pass)."""
    x = x.contiguous()
    Cc = x.shape[-1]
    npix = x.numel() // Cc
    out = torch.empty(2, Cc, dtype=torch.float64, device=x.device)
    lib().creste_chan_dot_workspace_bytes.restype = C.c_size_t
    n = 2 * lib().creste_chan_dot_workspace_bytes(C.c_longlong(npix), Cc)
    ws = _ws(n, x.device)
    check(lib().creste_chan_stats(ptr(x), C.c_longlong(npix), Cc, ptr(out), ptr(ws), C.c_size_t(n), stream()),
          "creste_chan_stats")
    return out


def maxpool2(x_nhwc):
    return maxpool2_concat([x_nhwc.contiguous()])


def maxpool2_bwd(x_nhwc, g):
    N, H, W, Cc = x_nhwc.shape
    dx = torch.empty_like(x_nhwc)
    check(lib().creste_maxpool2_bwd(ptr(x_nhwc.contiguous()), ptr(g.contiguous()), N, H, W, Cc,
                                    ptr(dx), stream()), "creste_maxpool2_bwd")
    return dx


def maxpool2_gather(x_nhwc, gg):
    N, H, W, Cc = x_nhwc.shape
    out = torch.empty(N, H // 2, W // 2, Cc, device=x_nhwc.device)
    check(lib().creste_maxpool2_gather(ptr(x_nhwc.contiguous()), ptr(gg.contiguous()), N, H, W, Cc,
                                       ptr(out), stream()), "creste_maxpool2_gather")
    return out


def upsample2(x_nhwc):
    """Bilinear x2 (align_corners=False), NHWC, C % 4 == 0."""
    N, H, W, _ = x_nhwc.shape
    return upsample_concat(None, x_nhwc.contiguous(), (2 * H, 2 * W), 2)


def upsample2_adjoint(g_nhwc):
    N, Ho, Wo, Cc = g_nhwc.shape
    Hi, Wi = Ho // 2, Wo // 2
    dx = torch.empty(N, Hi, Wi, Cc, device=g_nhwc.device)
    check(lib().creste_upsample_adjoint(ptr(g_nhwc.contiguous()), N, Hi, Wi, Cc, Ho, Wo,
                                        C.c_float(0.5), C.c_float(0.5), ptr(dx), stream()),
          "creste_upsample_adjoint")
    return dx


def upsample_adjoint(g_nhwc, Hi, Wi, ratio):
    """Adjoint of the bilinear up-sampling [N,Hi,Wi,C] -> [N,Ho,Wo,C] with sampling ratio `ratio` = 1 / scale."""
    N, Ho, Wo, Cc = g_nhwc.shape
    dx = torch.empty(N, Hi, Wi, Cc, device=g_nhwc.device)
    check(lib().creste_upsample_adjoint(ptr(g_nhwc.contiguous()), N, Hi, Wi, Cc, Ho, Wo, C.c_float(ratio),
                                        C.c_float(ratio), ptr(dx), stream()), "creste_upsample_adjoint")
    return dx


def upsample_adjoint_slice(g_nhwc, c0, Cc, Hi, Wi, ratio):
    """upsample_adjoint of the channel slice g[..., c0:c0+Cc], read in place (no chan_slice copy)."""
    g_nhwc = g_nhwc.contiguous()
    N, Ho, Wo, Cg = g_nhwc.shape
    dx = torch.empty(N, Hi, Wi, Cc, device=g_nhwc.device)
    check(lib().creste_upsample_adjoint_slice(ptr(g_nhwc), Cg, int(c0), N, Hi, Wi, int(Cc), Ho, Wo, C.c_float(ratio),
                                              C.c_float(ratio), ptr(dx), stream()), "creste_upsample_adjoint_slice")
    return dx


def conv2d_wgrad(x_nhwc, g_nhwc, R, S, pad):
    """dw [K,C,R,S] (torch layout) of a stride-1 conv: x [N,H,W,C], g [N,P,Q,K]."""
    x_nhwc, g_nhwc = x_nhwc.contiguous(), g_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    _, P, Q, K = g_nhwc.shape
    pt, pb, pl, pr = pad
    assert P == H + pt + pb - R + 1 and Q == W + pl + pr - S + 1
    d = ConvDesc(N, H, W, Cc, K, R, S, 1, pt, pl, P, Q, 0, 0, 0)
    lib().creste_conv2d_wgrad_workspace_bytes.restype = C.c_size_t
    n = lib().creste_conv2d_wgrad_workspace_bytes(C.byref(d))
    ws = _ws(n, x_nhwc.device)
    dw = torch.empty(R * S * Cc, K, device=x_nhwc.device)
    check(lib().creste_conv2d_wgrad(C.byref(d), ptr(x_nhwc), ptr(g_nhwc), ptr(dw), ptr(ws),
                                    C.c_size_t(n), stream()), "creste_conv2d_wgrad")
    return dw.view(R, S, Cc, K).permute(3, 2, 0, 1).contiguous()


def row_dot(x, y=None, mask=None):
    """out[b] = sum_i x[b,i] * y[b,i] * mask[b,i] over all trailing dims."""
    x = x.contiguous()
    B = x.shape[0]
    n = x.numel() // B
    out = torch.empty(B, device=x.device)
    check(lib().creste_row_dot(ptr(x), ptr(None if y is None else y.contiguous()),
                               ptr(None if mask is None else mask.contiguous()), B,
                               C.c_longlong(n), ptr(out), stream()), "creste_row_dot")
    return out


def row_scale(x, s, mask=None):
    x = x.contiguous()
    B = x.shape[0]
    out = torch.empty_like(x)
    check(lib().creste_row_scale(ptr(x), ptr(s.contiguous()),
                                 ptr(None if mask is None else mask.contiguous()), B,
                                 C.c_longlong(x.numel() // B), ptr(out), stream()), "creste_row_scale")
    return out


def row_normalize(x, mask=None, eps=1e-5):
    x = x.contiguous()
    B = x.shape[0]
    out = torch.empty_like(x)
    check(lib().creste_row_normalize(ptr(x), ptr(None if mask is None else mask.contiguous()), B,
                                     C.c_longlong(x.numel() // B), C.c_float(eps), ptr(out),
                                     stream()), "creste_row_normalize")
    return out


def grad_penalty(G_nchw):
    """mean over (b, pixel) of (||G[b,:,pixel]||_2 - 1)^2 -> 0-dim tensor."""
    G = G_nchw.contiguous()
    B, Cc = G.shape[0], G.shape[1]
    HW = G.numel() // (B * Cc)
    lib().creste_grad_penalty_workspace_bytes.restype = C.c_size_t
    n = lib().creste_grad_penalty_workspace_bytes(B, C.c_longlong(HW))
    ws = _ws(n, G.device)
    out = torch.empty((), device=G.device)
    check(lib().creste_grad_penalty(ptr(G), B, Cc, C.c_longlong(HW), ptr(out), ptr(ws),
                                    C.c_size_t(n), stream()), "creste_grad_penalty")
    return out


def grad_penalty_bwd(G_nchw, g_scalar):
    G = G_nchw.contiguous()
    B, Cc = G.shape[0], G.shape[1]
    HW = G.numel() // (B * Cc)
    dG = torch.empty_like(G)
    check(lib().creste_grad_penalty_bwd(ptr(G), ptr(g_scalar.contiguous().float()), B, Cc,
                                        C.c_longlong(HW), ptr(dG), stream()),
          "creste_grad_penalty_bwd")
    return dG


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """In-place torch.optim.Adam update of the flat fp32 buffers p, m, v from the flat gradient g."""
    check(lib().creste_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), C.c_longlong(p.numel()),
                                 C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                 int(step), C.c_float(grad_scale), stream()), "creste_adam_step")


def stage1_depth_losses(logits_nchw, pred_bins, label_mm, depth_min, depth_max, beta):
    """-> float64[4] device tensor {sum CE, #valid, #correct, sum smooth-L1} (stage-1 validation)."""
    logits = logits_nchw.contiguous().float()
    N, D = logits.shape[0], logits.shape[1]
    HW = logits.numel() // (N * D)
    bins = pred_bins.contiguous().to(torch.int64)
    lab = label_mm.contiguous().float()
    assert bins.numel() == N * HW and lab.numel() == N * HW
    acc = torch.empty(4, dtype=torch.float64, device=logits.device)
    check(lib().creste_stage1_depth_losses(ptr(logits), ptr(bins), ptr(lab), N, D, C.c_longlong(HW),
                                           C.c_float(depth_min), C.c_float(depth_max), C.c_float(beta),
                                           ptr(acc), stream()), "creste_stage1_depth_losses")
    return acc


def masked_mse(pred, gt):
    """-> float64[2] device tensor {sum of squared differences where gt is not +-inf, count}."""
    pred, gt = pred.contiguous().float(), gt.contiguous().float()
    assert pred.numel() == gt.numel()
    acc = torch.empty(2, dtype=torch.float64, device=pred.device)
    check(lib().creste_masked_mse(ptr(pred), ptr(gt), C.c_longlong(pred.numel()), ptr(acc), stream()),
          "creste_masked_mse")
    return acc


# ------------------------------------------------------- stage-1 backbone training primitives
def _npix_c(x):
    Cc = x.shape[-1]
    return x.numel() // Cc, Cc


def chan_moments(x):
    """-> float64 [2, C]: per-channel (sum x, sum x^2) over all leading dims; any C % 4 == 0."""
    x = x.contiguous()
    npix, Cc = _npix_c(x)
    out = torch.empty(2, Cc, dtype=torch.float64, device=x.device)
    n = lib().creste_chan_reduce_workspace_bytes(C.c_longlong(npix), Cc, 2, 1)
    ws = _ws(n, x.device)
    check(lib().creste_chan_moments(ptr(x), C.c_longlong(npix), Cc, ptr(out), ptr(ws), C.c_size_t(n),
                                    stream()), "creste_chan_moments")
    return out


USE_PUBLISHED_AMAX = os.environ.get("CRESTE_NO_PUBLISHED_AMAX") is None     # experiment / test switch
PUBLISH_AMAX_MODE = None       # set by engine.set_precision: producers measure max|out| only in the 3xFP16 mode


def _want_amax(t):
    """The generic elementwise producers (chan_affine, relu_bwd) measure max|out| when a tensor-core conv may consume
    the result: 3xFP16 mode, a 4-D activation with a channel count the tensor-core path takes."""
    return PUBLISH_AMAX_MODE == "3xfp16" and USE_PUBLISHED_AMAX and t.dim() == 4 and t.shape[-1] % 8 == 0


def publish_amax(t, amax):
    """Attach the exact max|t| (device float[1]) the producing kernel measured; split_f16 then skips its amax pass.
    Valid while the tensor is not written again (checked through torch's version counter) and within the same
    capture state (a value measured by an eager warm-up is not consumed inside a CUDA-graph capture)."""
    t._amax_exact = (amax, t._version, t.is_cuda and torch.cuda.is_current_stream_capturing())
    return t


def published_amax(t):
    rec = getattr(t, "_amax_exact", None)
    if rec is None or rec[1] != t._version:
        return None
    return rec[0] if rec[2] == (t.is_cuda and torch.cuda.is_current_stream_capturing()) else None


def chan_affine_act(x, a, b, act, want_amax=False):
    """y = act(x * a[c] + b[c]), act in {'none', 'relu', 'swish'}.  want_amax: the kernel also measures max|y| and
    the result carries it (publish_amax) for the 3xFP16 operand split of the conv that consumes y."""
    x = x.contiguous()
    npix, Cc = _npix_c(x)
    y = torch.empty_like(x)
    if want_amax:
        amax = torch.empty(1, device=x.device)
        check(lib().creste_chan_affine_act_amax(ptr(x), ptr(a), ptr(b), C.c_longlong(npix), Cc, ACT[act], ptr(y),
                                                ptr(amax), stream()), "creste_chan_affine_act_amax")
        return publish_amax(y, amax)
    check(lib().creste_chan_affine_act(ptr(x), ptr(a), ptr(b), C.c_longlong(npix), Cc, ACT[act], ptr(y),
                                       stream()), "creste_chan_affine_act")
    return y


def bn_act_bwd(g, x, a, b, act, want_gu=True):
    """-> (gu, sums float64 [2, C]) with gu = g * act'(x*a+b) (gu is g itself for act 'none') and
    sums = (sum gu, sum gu*x) per channel.  want_gu=False: gu is not stored (None is returned for act != 'none';
    chan_axpby_act recomputes it)."""
    g, x = g.contiguous(), x.contiguous()
    npix, Cc = _npix_c(x)
    gu = torch.empty_like(g) if (ACT[act] != 0 and want_gu) else None
    sums = torch.empty(2, Cc, dtype=torch.float64, device=x.device)
    n = lib().creste_chan_reduce_workspace_bytes(C.c_longlong(npix), Cc, 2, 1)
    ws = _ws(n, x.device)
    check(lib().creste_bn_act_bwd(ptr(g), ptr(x), ptr(a), ptr(b), C.c_longlong(npix), Cc, ACT[act], ptr(gu),
                                  ptr(sums), ptr(ws), C.c_size_t(n), stream()), "creste_bn_act_bwd")
    return (gu if (gu is not None or ACT[act] != 0) else g), sums


def chan_axpby_act(g, x, a, b, act, p, q, r, want_amax=False):
    """out = (g * act'(x*a[c]+b[c])) * p[c] + x*q[c] + r[c]: chan_axpby on the gu that bn_act_bwd would have stored."""
    g, x = g.contiguous(), x.contiguous()
    npix, Cc = _npix_c(x)
    out = torch.empty_like(x)
    amax = torch.empty(1, device=x.device) if want_amax else None
    check(lib().creste_chan_axpby_act(ptr(g), ptr(x), ptr(a.contiguous()), ptr(b.contiguous()), ACT[act],
                                      ptr(p.contiguous()), ptr(q.contiguous()), ptr(r.contiguous()), C.c_longlong(npix),
                                      Cc, ptr(out), ptr(amax), stream()), "creste_chan_axpby_act")
    return publish_amax(out, amax) if want_amax else out


def chan_axpby(u, x, p, q, r, want_amax=False):
    """out = u*p[c] + x*q[c] + r[c].  want_amax: as in chan_affine_act."""
    u, x = u.contiguous(), x.contiguous()
    npix, Cc = _npix_c(x)
    out = torch.empty_like(x)
    if want_amax:
        amax = torch.empty(1, device=x.device)
        check(lib().creste_chan_axpby_amax(ptr(u), ptr(x), ptr(p.contiguous()), ptr(q.contiguous()), ptr(r.contiguous()),
                                           C.c_longlong(npix), Cc, ptr(out), ptr(amax), stream()),
              "creste_chan_axpby_amax")
        return publish_amax(out, amax)
    check(lib().creste_chan_axpby(ptr(u), ptr(x), ptr(p.contiguous()), ptr(q.contiguous()), ptr(r.contiguous()),
                                  C.c_longlong(npix), Cc, ptr(out), stream()), "creste_chan_axpby")
    return out


def _dw_out(H, W, R, stride, pad):
    pt, pb, pl, pr = pad
    return (H + pt + pb - R) // stride + 1, (W + pl + pr - R) // stride + 1


def dwconv_fwd(x_nhwc, w_rsc, R, stride, pad):
    """Raw depthwise conv: x [N,H,W,C], w [R*R, C], pad = (top, bottom, left, right)."""
    x_nhwc = x_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    P, Q = _dw_out(H, W, R, stride, pad)
    y = torch.empty(N, P, Q, Cc, device=x_nhwc.device)
    check(lib().creste_dwconv_fwd(ptr(x_nhwc), ptr(w_rsc.contiguous()), N, H, W, Cc, R, stride, pad[0], pad[2],
                                  P, Q, ptr(y), stream()), "creste_dwconv_fwd")
    return y


def dwconv_dgrad(g_nhwc, w_rsc, x_shape, R, stride, pad):
    g_nhwc = g_nhwc.contiguous()
    N, H, W, Cc = x_shape
    _, P, Q, _ = g_nhwc.shape
    dx = torch.empty(N, H, W, Cc, device=g_nhwc.device)
    check(lib().creste_dwconv_dgrad(ptr(g_nhwc), ptr(w_rsc.contiguous()), N, H, W, Cc, R, stride, pad[0],
                                    pad[2], P, Q, ptr(dx), stream()), "creste_dwconv_dgrad")
    return dx


def dwconv_wgrad(x_nhwc, g_nhwc, R, stride, pad):
    """-> dw [R*R, C]."""
    x_nhwc, g_nhwc = x_nhwc.contiguous(), g_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    _, P, Q, _ = g_nhwc.shape
    dw = torch.empty(R * R, Cc, device=x_nhwc.device)
    n = lib().creste_dwconv_wgrad_workspace_bytes(N, Cc, R, P, Q)
    ws = _ws(n, x_nhwc.device)
    check(lib().creste_dwconv_wgrad(ptr(x_nhwc), ptr(g_nhwc), N, H, W, Cc, R, stride, pad[0], pad[2], P, Q,
                                    ptr(dw), ptr(ws), C.c_size_t(n), stream()), "creste_dwconv_wgrad")
    return dw


def sample_dot(x, y=None, scale=1.0):
    """out[b, c] = scale * sum_pix x[b,pix,c] * y[b,pix,c]  (y None: plain sum) -> [B,1,1,C]."""
    x = x.contiguous()
    y = None if y is None else y.contiguous()
    B, Cc = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * Cc)
    out = torch.empty(B, 1, 1, Cc, device=x.device)
    n = lib().creste_chan_reduce_workspace_bytes(C.c_longlong(HW), Cc, 1, B)
    ws = _ws(n, x.device)
    check(lib().creste_sample_dot(ptr(x), ptr(y), B, C.c_longlong(HW), Cc, C.c_float(scale), ptr(out), ptr(ws),
                                  C.c_size_t(n), stream()), "creste_sample_dot")
    return out


def sample_affine(x, a=None, b=None, shape=None):
    """out[b,pix,c] = (x * a[b,c] if x is not None else 0) + b[b,c]; `shape` when x is None."""
    shape = tuple(x.shape) if x is not None else tuple(shape)
    B, Cc = shape[0], shape[-1]
    HW = 1
    for d in shape[1:-1]:
        HW *= d
    dev = (x if x is not None else b).device
    out = torch.empty(shape, device=dev)
    check(lib().creste_sample_affine(ptr(None if x is None else x.contiguous()),
                                     ptr(None if a is None else a.contiguous()),
                                     ptr(None if b is None else b.contiguous()), B, C.c_longlong(HW), Cc,
                                     ptr(out), stream()), "creste_sample_affine")
    return out


def act(x, kind):
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib().creste_act(ptr(x), C.c_longlong(x.numel()), ACT[kind], ptr(y), stream()), "creste_act")
    return y


def act_bwd(g, x, kind):
    g, x = g.contiguous(), x.contiguous()
    dx = torch.empty_like(x)
    check(lib().creste_act_bwd(ptr(g), ptr(x), C.c_longlong(x.numel()), ACT[kind], ptr(dx), stream()),
          "creste_act_bwd")
    return dx


def add_scaled(inp, x, s=None):
    """out = inp + x * s[b]  (s None: plain sum) -- identity skip with drop-connect."""
    inp, x = inp.contiguous(), x.contiguous()
    B = x.shape[0]
    out = torch.empty_like(x)
    check(lib().creste_add_scaled(ptr(inp), ptr(x), ptr(None if s is None else s.contiguous()), B,
                                  C.c_longlong(x.numel() // B), ptr(out), stream()), "creste_add_scaled")
    return out


def chan_slice(x, c0, cn):
    x = x.contiguous()
    npix, Cc = _npix_c(x)
    out = torch.empty(*x.shape[:-1], cn, device=x.device)
    check(lib().creste_chan_slice(ptr(x), C.c_longlong(npix), Cc, int(c0), int(cn), ptr(out), stream()),
          "creste_chan_slice")
    return out


def wgrad_strided(x_nhwc, g_nhwc, R, S, stride, pad):
    """dw [K,C,R,S] (torch layout) of a strided dense conv with C == 4 (the EfficientNet stem)."""
    x_nhwc, g_nhwc = x_nhwc.contiguous(), g_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    _, P, Q, K = g_nhwc.shape
    n = lib().creste_wgrad_strided_workspace_bytes(N, P, Q, Cc, K, R, S)
    ws = _ws(n, x_nhwc.device)
    dw = torch.empty(R * S * Cc, K, device=x_nhwc.device)
    check(lib().creste_wgrad_strided(ptr(x_nhwc), ptr(g_nhwc), N, H, W, Cc, K, R, S, stride, pad[0], pad[2],
                                     P, Q, ptr(dw), ptr(ws), C.c_size_t(n), stream()), "creste_wgrad_strided")
    return dw.view(R, S, Cc, K).permute(3, 2, 0, 1).contiguous()


def ce_depth_bwd(logits_nchw, label_mm, depth_min, depth_max, scale_dev):
    logits = logits_nchw.contiguous()
    N, D = logits.shape[0], logits.shape[1]
    HW = logits.numel() // (N * D)
    out = torch.empty_like(logits)
    check(lib().creste_ce_depth_bwd(ptr(logits), ptr(label_mm.contiguous().float()), N, D, C.c_longlong(HW),
                                    C.c_float(depth_min), C.c_float(depth_max),
                                    ptr(scale_dev.reshape(1).float().contiguous()), ptr(out), stream()),
          "creste_ce_depth_bwd")
    return out


def masked_mse_bwd(pred, gt, scale_dev):
    pred, gt = pred.contiguous(), gt.contiguous().float()
    out = torch.empty_like(pred)
    check(lib().creste_masked_mse_bwd(ptr(pred), ptr(gt), C.c_longlong(pred.numel()),
                                      ptr(scale_dev.reshape(1).float().contiguous()), ptr(out), stream()),
          "creste_masked_mse_bwd")
    return out


def wgrad_tc_supported(x_shape, K, R, S, pad):
    N, H, W, Cc = x_shape
    pt, pb, pl, pr = pad
    d = ConvDesc(N, H, W, Cc, K, R, S, 1, pt, pl, H + pt + pb - R + 1, W + pl + pr - S + 1, 0, 0, 4)
    return bool(lib().creste_conv2d_wgrad_tc_supported(C.byref(d)))


def conv2d_wgrad_tc(x_nhwc, g_nhwc, R, S, pad):
    """dw [K,C,R,S] of a stride-1 conv on the tcgen05 tensor cores (3xFP16 split); C, K >= 64."""
    x_nhwc, g_nhwc = x_nhwc.contiguous(), g_nhwc.contiguous()
    N, H, W, Cc = x_nhwc.shape
    _, P, Q, K = g_nhwc.shape
    pt, pb, pl, pr = pad
    assert P == H + pt + pb - R + 1 and Q == W + pl + pr - S + 1
    d = ConvDesc(N, H, W, Cc, K, R, S, 1, pt, pl, P, Q, 0, 0, 4)
    n = lib().creste_conv2d_wgrad_tc_workspace_bytes(C.byref(d))
    ws = _ws(n, x_nhwc.device)
    dw = torch.empty(R * S * Cc, K, device=x_nhwc.device)
    check(lib().creste_conv2d_wgrad_tc(C.byref(d), ptr(x_nhwc), ptr(g_nhwc), ptr(dw), ptr(ws), C.c_size_t(n),
                                       stream()), "creste_conv2d_wgrad_tc")
    return dw.view(R, S, Cc, K).permute(3, 2, 0, 1).contiguous()


def split_f16(x_nhwc):
    """The 3xFP16 operand of a dense fp32 NHWC tensor (amax -> power-of-two scale -> fp16 hi / lo) as a SplitAct."""
    amax = published_amax(x_nhwc)
    x_nhwc = x_nhwc.contiguous()
    hi = torch.empty(x_nhwc.shape, dtype=torch.float16, device=x_nhwc.device)
    lo = torch.empty_like(hi)
    scal = torch.empty(4, device=x_nhwc.device)
    if amax is not None and USE_PUBLISHED_AMAX:     # measured by the producer: one pass, same scale, same halves
        check(lib().creste_f16_split_amax(ptr(x_nhwc), C.c_longlong(x_nhwc.numel()), ptr(amax), ptr(hi), ptr(lo),
                                          ptr(scal), stream()), "creste_f16_split_amax")
        return SplitAct(hi, lo, scal, x_nhwc.shape)
    check(lib().creste_f16_split(ptr(x_nhwc), C.c_longlong(x_nhwc.numel()), ptr(hi), ptr(lo), ptr(scal), stream()),
          "creste_f16_split")
    return SplitAct(hi, lo, scal, x_nhwc.shape)


def split_f16_cached(x_nhwc):
    """split_f16 remembered on the tensor (until it is written again): in the stage-3 graph the same activation /
    gradient is the operand of up to four tensor-core convs (forward conv, weight gradient, and their counterparts in
    the double backward of the gradient penalty), in stage 1 the decoder features feed two heads.
    A record made outside a CUDA-graph capture is never used inside one (and vice versa): a split of a static input
    computed by the eager warm-up would otherwise be missing from the captured graph and go stale on replay."""
    capturing = x_nhwc.is_cuda and torch.cuda.is_current_stream_capturing()
    rec = getattr(x_nhwc, "_split_rec", None)
    if rec is not None and rec[1] == x_nhwc._version and rec[2] == capturing and USE_PUBLISHED_AMAX:
        return rec[0]
    sa = split_f16(x_nhwc)
    if x_nhwc.is_contiguous():
        x_nhwc._split_rec = (sa, x_nhwc._version, capturing)
    return sa


def conv2d_wgrad_tc_presplit(xs, gs, R, S, pad):
    """conv2d_wgrad_tc on SplitAct operands (x: the forward conv's saved operand, g: the split output gradient)."""
    N, H, W, Cc = xs.shape
    _, P, Q, K = gs.shape
    pt, pb, pl, pr = pad
    assert P == H + pt + pb - R + 1 and Q == W + pl + pr - S + 1
    d = ConvDesc(N, H, W, Cc, K, R, S, 1, pt, pl, P, Q, 0, 0, 4)
    n = lib().creste_conv2d_wgrad_tc_workspace_bytes(C.byref(d))
    ws = _ws(n, xs.device)
    dw = torch.empty(R * S * Cc, K, device=xs.device)
    check(lib().creste_conv2d_wgrad_tc_presplit(C.byref(d), ptr(xs.hi), ptr(xs.lo), ptr(xs.scal), ptr(gs.hi), ptr(gs.lo),
                                                ptr(gs.scal), ptr(dw), ptr(ws), C.c_size_t(n), stream()),
          "creste_conv2d_wgrad_tc_presplit")
    return dw.view(R, S, Cc, K).permute(3, 2, 0, 1).contiguous()


def wgrad_rows(x, g):
    """dw [K,C,1,1] of a 1x1 conv over a handful of rows: x [...,C], g [...,K] with <= 4096 rows."""
    x, g = x.contiguous(), g.contiguous()
    Cc, K = x.shape[-1], g.shape[-1]
    npix = x.numel() // Cc
    dw = torch.empty(Cc, K, device=x.device)
    check(lib().creste_wgrad_rows(ptr(x), ptr(g), npix, Cc, K, ptr(dw), stream()), "creste_wgrad_rows")
    return dw.t().contiguous().view(K, Cc, 1, 1)


def bn_fwd_finalize(stats, weight, bias, M, eps, momentum, running_mean, running_var):
    """moments [2,C] float64 -> (ab [2,C] float = scale / shift of the normalisation, mean_inv [2,C] float64);
    updates running_mean / running_var in place when given (momentum form of F.batch_norm)."""
    Cc = stats.shape[1]
    ab = torch.empty(2, Cc, device=stats.device)
    mi = torch.empty(2, Cc, dtype=torch.float64, device=stats.device)
    check(lib().creste_bn_fwd_finalize(ptr(stats.contiguous(), torch.float64), ptr(weight), ptr(bias), Cc,
                                       C.c_double(float(M)), C.c_double(float(eps)), C.c_float(float(momentum or 0.0)),
                                       ptr(running_mean), ptr(running_var), ptr(ab), ptr(mi), stream()),
          "creste_bn_fwd_finalize")
    return ab, mi


def bn_bwd_finalize(sums, ab, mean_inv, M):
    """(sum gu, sum gu*x) -> [4,C] float = (dgamma, dbeta, q, r) with dx = gu*a + x*q + r."""
    Cc = sums.shape[1]
    out = torch.empty(4, Cc, device=sums.device)
    check(lib().creste_bn_bwd_finalize(ptr(sums.contiguous(), torch.float64), ptr(ab), ptr(mean_inv), Cc,
                                       C.c_double(float(M)), ptr(out), stream()), "creste_bn_bwd_finalize")
    return out


def pack_conv_weight_f16_strided(w):
    """pack_conv_weight_f16 in one kernel launch, reading w [K,C,R,S] through its strides (no copy for the
    transposed view the data-gradient conv passes).  Same layout; the scale is taken from the exponent bits."""
    assert w.dtype == torch.float32 and w.is_cuda
    K, Cc, R, S = w.shape
    _, npad, _ = tc_layout(K, Cc, R, S)
    cpad = (Cc + 63) // 64 * 64
    out = torch.empty(npad * R * S * cpad + npad, device=w.device)
    sK, sC, sR, sS = w.stride()
    check(lib().creste_pack_weight_f16(C.c_void_p(w.data_ptr()), C.c_longlong(sK), C.c_longlong(sC), C.c_longlong(sR),
                                       C.c_longlong(sS), K, Cc, R, S, ptr(out), stream()), "creste_pack_weight_f16")
    return out


# ------------------------------------------------------------------------- stage-2 loss kernels
def _mask_ptr(mask):
    return ptr(None if mask is None else mask.contiguous().to(torch.uint8))


def smooth_l1(pred, gt, mask, gt_scale, beta):
    """-> float64[2] device tensor {sum of smooth-L1(pred - gt * gt_scale), count} over mask & isfinite(gt)."""
    pred, gt = pred.contiguous().float(), gt.contiguous().float()
    assert pred.numel() == gt.numel()
    acc = torch.empty(2, dtype=torch.float64, device=pred.device)
    check(lib().creste_smooth_l1(ptr(pred), ptr(gt), _mask_ptr(mask), C.c_longlong(pred.numel()), C.c_float(gt_scale),
                                 C.c_float(beta), ptr(acc), stream()), "creste_smooth_l1")
    return acc


def smooth_l1_bwd(pred, gt, mask, gt_scale, beta, scale_dev):
    pred, gt = pred.contiguous().float(), gt.contiguous().float()
    out = torch.empty_like(pred)
    check(lib().creste_smooth_l1_bwd(ptr(pred), ptr(gt), _mask_ptr(mask), C.c_longlong(pred.numel()), C.c_float(gt_scale),
                                     C.c_float(beta), ptr(scale_dev.reshape(1).float().contiguous()), ptr(out), stream()),
          "creste_smooth_l1_bwd")
    return out


def ce_weighted(logits_nchw, labels, mask, class_weights, ignore_index=-100):
    """-> float64[4] {sum w*nll, sum w, #correct over label != 0, #(label != 0)} over the masked cells."""
    logits = logits_nchw.contiguous().float()
    B, Cc = logits.shape[0], logits.shape[1]
    HW = logits.numel() // (B * Cc)
    acc = torch.empty(4, dtype=torch.float64, device=logits.device)
    check(lib().creste_ce_weighted(ptr(logits), ptr(labels.contiguous().to(torch.int64)), _mask_ptr(mask),
                                   ptr(None if class_weights is None else class_weights.contiguous().float()), B, Cc,
                                   C.c_longlong(HW), C.c_longlong(int(ignore_index)), ptr(acc), stream()),
          "creste_ce_weighted")
    return acc


def ce_weighted_bwd(logits_nchw, labels, mask, class_weights, ignore_index, scale_dev):
    logits = logits_nchw.contiguous().float()
    B, Cc = logits.shape[0], logits.shape[1]
    HW = logits.numel() // (B * Cc)
    out = torch.empty_like(logits)
    check(lib().creste_ce_weighted_bwd(ptr(logits), ptr(labels.contiguous().to(torch.int64)), _mask_ptr(mask),
                                       ptr(None if class_weights is None else class_weights.contiguous().float()), B, Cc,
                                       C.c_longlong(HW), C.c_longlong(int(ignore_index)),
                                       ptr(scale_dev.reshape(1).float().contiguous()), ptr(out), stream()),
          "creste_ce_weighted_bwd")
    return out


def l2norm_rows(x, eps=1e-12):
    x = x.contiguous().float()
    N, D = x.shape
    y, nrm = torch.empty_like(x), torch.empty(N, device=x.device)
    check(lib().creste_l2norm_rows(ptr(x), N, D, C.c_float(eps), ptr(y), ptr(nrm), stream()), "creste_l2norm_rows")
    return y, nrm


def l2norm_rows_bwd(y, dy, nrm):
    N, D = y.shape
    dx = torch.empty_like(y)
    check(lib().creste_l2norm_rows_bwd(ptr(y.contiguous()), ptr(dy.contiguous().float()), ptr(nrm), N, D, ptr(dx),
                                       stream()), "creste_l2norm_rows_bwd")
    return dx


def supcon_fwd(f, a, lf, la, self_off, temperature, class_weights=None):
    """-> (stats [N,4], loss_sum float64[1]); f [N,D] / a [Na,D] L2-normalised, lf / la int64 labels."""
    f, a = f.contiguous().float(), a.contiguous().float()
    N, D = f.shape
    stats = torch.empty(N, 4, device=f.device)
    acc = torch.empty(1, dtype=torch.float64, device=f.device)
    check(lib().creste_supcon_fwd(ptr(f), ptr(a), ptr(lf.contiguous().to(torch.int64)), ptr(la.contiguous().to(torch.int64)),
                                  N, a.shape[0], D, int(self_off), C.c_float(temperature),
                                  ptr(None if class_weights is None else class_weights.contiguous().float()), ptr(stats),
                                  ptr(acc), stream()), "creste_supcon_fwd")
    return stats, acc


def supcon_bwd(f, a, lf, la, self_off, temperature, class_weights, stats, scale_dev):
    f, a = f.contiguous().float(), a.contiguous().float()
    N, D = f.shape
    df, da = torch.empty_like(f), torch.empty_like(a)
    check(lib().creste_supcon_bwd(ptr(f), ptr(a), ptr(lf.contiguous().to(torch.int64)), ptr(la.contiguous().to(torch.int64)),
                                  N, a.shape[0], D, int(self_off), C.c_float(temperature),
                                  ptr(None if class_weights is None else class_weights.contiguous().float()),
                                  ptr(stats.contiguous()), ptr(scale_dev.reshape(1).float().contiguous()), ptr(df), ptr(da),
                                  stream()), "creste_supcon_bwd")
    return df, da


# ------------------------------------------------------------------ export path (torch.jit.trace)
# Under torch.jit.trace the eval-forward entry points route through the `creste::` dispatcher ops of
# creste_public_b200.torch_ops so that the tracer records them (scripts/runtime/compile.py:197 of the reference);
# everywhere else the functions above are called directly.
_RAW = {n: globals()[n] for n in ("conv2d", "dwconv_bn_swish", "se_gate", "upsample_concat", "maxpool2_concat",
                                  "nchw_to_nhwc", "nhwc_to_nchw", "depth_expectation", "frustum_to_bev", "zmlp_concat",
                                  "splat_soft", "proj_head")}


def _tracing():
    if not torch.jit.is_tracing():
        return False
    from . import torch_ops
    return not torch_ops.inside()


def _t_conv2d(x_nhwc, w_packed, K, R, S, stride=1, pad=(0, 0, 0, 0), scale=None, shift=None, gate=None, residual=None,
              act="none", out_nchw=False, precision="fp32", amax_in=None, amax_out=None, split_out=None):
    if _tracing():
        return torch.ops.creste.conv2d(x_nhwc, w_packed, int(K), int(R), int(S), int(stride), [int(p) for p in pad], scale,
                                       shift, gate, residual, act or "none", bool(out_nchw), precision)
    return _RAW["conv2d"](x_nhwc, w_packed, K, R, S, stride, pad, scale, shift, gate, residual, act, out_nchw, precision,
                          amax_in, amax_out, split_out)


def _t_dwconv_bn_swish(x_nhwc, w_rsc, scale, shift, R, stride, pad, amax_out=None):
    if _tracing():
        return torch.ops.creste.dwconv_bn_swish(x_nhwc, w_rsc, scale, shift, int(R), int(stride), [int(p) for p in pad])
    return _RAW["dwconv_bn_swish"](x_nhwc, w_rsc, scale, shift, R, stride, pad, amax_out)


def _t_se_gate(chan_part, hw, w_red, b_red, w_exp, b_exp):
    if _tracing():
        return torch.ops.creste.se_gate(chan_part, int(hw), w_red, b_red, w_exp, b_exp)
    return _RAW["se_gate"](chan_part, hw, w_red, b_red, w_exp, b_exp)


def _t_upsample_concat(skip_nhwc, x_nhwc, out_hw, scale_factor=None, x_first=False):
    if _tracing():
        Hi, Wi = x_nhwc.shape[1], x_nhwc.shape[2]
        if scale_factor is None:
            rh, rw = Hi / out_hw[0], Wi / out_hw[1]
        else:
            sh, sw = (scale_factor, scale_factor) if not isinstance(scale_factor, (tuple, list)) else scale_factor
            rh, rw = 1.0 / sh, 1.0 / sw
        return torch.ops.creste.upsample_concat(skip_nhwc, x_nhwc, [int(out_hw[0]), int(out_hw[1])], float(rh), float(rw),
                                                bool(x_first))
    return _RAW["upsample_concat"](skip_nhwc, x_nhwc, out_hw, scale_factor, x_first)


def _t_maxpool2_concat(srcs_nhwc, rows_out=None, want_nchw=False):
    if _tracing():
        rows = srcs_nhwc[0].shape[1] // 2 if rows_out is None else rows_out
        a, b = torch.ops.creste.maxpool2_concat(list(srcs_nhwc), int(rows))
        return (a, b) if want_nchw else a
    return _RAW["maxpool2_concat"](srcs_nhwc, rows_out, want_nchw)


def _t_nchw_to_nhwc(x):
    return torch.ops.creste.nchw_to_nhwc(x.contiguous()) if _tracing() else _RAW["nchw_to_nhwc"](x)


def _t_nhwc_to_nchw(x):
    return torch.ops.creste.nhwc_to_nchw(x.contiguous()) if _tracing() else _RAW["nhwc_to_nchw"](x)


def _t_depth_expectation(logits_nhwc, dmin=300.0, dmax=25600.0, out_div=1000.0):
    if _tracing():
        return torch.ops.creste.depth_expectation(logits_nhwc, float(dmin), float(dmax), float(out_div))
    return _RAW["depth_expectation"](logits_nhwc, dmin, dmax, out_div)


def _t_frustum_to_bev(depth, p2p, pc_range, voxel):
    if _tracing():
        return torch.ops.creste.frustum_to_bev(depth.contiguous().float(), p2p.contiguous().float(),
                                               [float(v) for v in pc_range], [float(voxel[0]), float(voxel[1])])
    return _RAW["frustum_to_bev"](depth, p2p, pc_range, voxel)


def _t_zmlp_concat(feats_nhwc, z, w1, b1, w2, b2, amax_out=None):
    if _tracing():
        return torch.ops.creste.zmlp_concat(feats_nhwc, z, w1, b1, w2, b2)
    return _RAW["zmlp_concat"](feats_nhwc, z, w1, b1, w2, b2, amax_out)


def _t_splat_soft(xy, feats_nhwc, mask, H, W, min_weight=1.0, want_nhwc=True, want_nchw=True, want_idx=False):
    if _tracing() and feats_nhwc is not None and not want_idx:
        a, b, d = torch.ops.creste.splat_soft(xy, feats_nhwc, mask, int(H), int(W), float(min_weight))
        return {"bev_nhwc": a if want_nhwc else None, "bev_nchw": b if want_nchw else None, "dens": d, "idx": None}
    return _RAW["splat_soft"](xy, feats_nhwc, mask, H, W, min_weight, want_nhwc, want_nchw, want_idx)


def _t_proj_head(x_nhwc, w_kc, bias, want_nchw=True, want_x_nchw=True):
    if _tracing():
        a, b, c = torch.ops.creste.proj_head(x_nhwc, w_kc, bias)
        return a, (b if want_nchw else None), (c if want_x_nchw else None)
    return _RAW["proj_head"](x_nhwc, w_kc, bias, want_nchw, want_x_nchw)


for _n in list(_RAW):
    _w = globals()["_t_" + _n]
    _w.__doc__ = _RAW[_n].__doc__
    _w.__name__ = _n
    globals()[_n] = _w
del _n, _w
