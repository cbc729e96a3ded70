"""Differentiable wrappers of the sm_100a kernels for the stage-3 (IRL) training step.

The reference trains the reward FCN with PyTorch autograd, including a *double backward*: the
SMODICE gradient penalty differentiates d(sum r)/d(input_view) w.r.t. the weights
(creste/utils/loss_utils.py:1208-1217).  torch.autograd is used here as the tape engine only:
every node is a `torch.autograd.Function` whose forward is one C-ABI kernel and whose backward is
written *in terms of the other Functions in this file*, so the set is closed under
differentiation and `torch.autograd.grad(..., create_graph=True)` works to any order:

    Conv2dFn    <-> Conv2dFn with flipped-transposed weights (dgrad), WGradFn
    WGradFn     <-> Conv2dFn (both arguments)
    ChanAffineFn (+ReLU) <-> ReluBwdFn, ChanDotFn ;  ChanDotFn <-> ChanAffineFn
    MaxPool2Fn  <-> MaxPoolBwdFn <-> MaxPoolGatherFn
    Up2Fn       <-> Up2AdjFn ;  ToNHWC <-> ToNCHW ;  RowDotFn <-> RowScaleFn
    GradPenaltyFn (first order only: it is the last node before the loss)

Tensors are channels-last fp32.  The only torch-native differentiable ops in the graph are
views / permutations / pads of tiny tensors (weights, per-channel vectors) and `torch.cat`.
"""
import os

import torch
from torch.autograd import Function

from . import engine, ops


def _pad_last(t, mult=4):
    c = t.shape[-1]
    r = (-c) % mult
    if r == 0:
        return t.contiguous()
    return torch.cat([t, t.new_zeros(*t.shape[:-1], r)], dim=-1).contiguous()


def _pad4(ph, pw):
    """(ph, pw) with each an int or a (low, high) pair -> (top, bottom, left, right)."""
    t, b = (ph, ph) if isinstance(ph, int) else ph
    l, r = (pw, pw) if isinstance(pw, int) else pw
    return (int(t), int(b), int(l), int(r))


def _conv_raw(x, w, ph, pw, stride=1):
    """Conv of NHWC x with torch-layout weights w [K,C,R,S]; no epilogue.  Channel counts that are not
    multiples of 4 are zero-padded (exact).  ph / pw: symmetric int or (low, high) pair."""
    K, Cc, R, S = w.shape
    xp = _pad_last(x)
    Cp = xp.shape[-1]
    Kp = K + ((-K) % 4)
    wp = w
    if Cp != Cc or Kp != K:
        wp = w.new_zeros(Kp, Cp, R, S)
        wp[:K, :Cc] = w
    pad = _pad4(ph, pw)
    mode = engine.pick_mode(tuple(xp.shape), Kp, R, S, stride, pad, engine.get_precision())
    if xp.shape[1] * xp.shape[2] < 16:
        mode = "fp32"      # squeeze-excite vectors [B,1,1,C]: a handful of rows, no tensor-core tile
    if mode == "fp32":
        packed = ops.pack_conv_weight(wp.detach().float())
    elif mode == "3xfp16" and PRESPLIT_REUSE and stride == 1 and Cp == Cc and Kp == K and Cc % 8 == 0 and K % 8 == 0:
        # the operand split is remembered on the tensor (and takes the producer's published maximum): the same x
        # is split once for this conv, its weight gradient and their double-backward counterparts
        packed = ops.pack_conv_weight_f16_strided(wp.detach().float())
        return ops.conv2d_presplit(ops.split_f16_cached(xp), packed, K, R, S, 1, pad, precision="3xfp16")
    elif mode in ("3xfp16", "fp16"):
        packed = ops.pack_conv_weight_f16_strided(wp.detach().float())      # one launch, strided read
    else:
        packed = ops.pack_conv_weight_tc(wp.detach().float(), split=(mode == "3xtf32"))
    y = ops.conv2d(xp, packed, Kp, R, S, stride, pad, precision=mode)
    return y if Kp == K else y[..., :K].contiguous()


WGRAD_TC = True      # tcgen05 weight gradient (>= 512 output pixels) in the tensor-core precision modes
WGRAD_TILE = 64      # the CUDA-core wgrad kernels hold one [C x K] tile of at most 64 x 64 per CTA


def _wgrad_raw(x, g, R, S, ph, pw):
    """dw [K,C,R,S] of a stride-1 conv.  Layers wider than 64 channels (the backbone's) are cut into
    64-channel slices of x and g (one extra pass over each tensor) and every (c, k) tile pair goes
    through the same kernel."""
    Cc, K = x.shape[-1], g.shape[-1]
    pad = _pad4(ph, pw)
    if R == 1 and S == 1 and x.numel() // Cc <= 64 and max(Cc, K) > WGRAD_TILE:
        return ops.wgrad_rows(x, g)          # squeeze-excite convs: [B,1,1,C] vectors, one launch
    if WGRAD_TC and engine.get_precision() != "fp32" and x.dim() == 4:
        Cp, Kp = Cc + (-Cc) % 8, K + (-K) % 8
        if ops.wgrad_tc_supported(tuple(x.shape[:3]) + (Cp,), Kp, R, S, pad):      # tcgen05, 3xFP16 split
            if PRESPLIT_REUSE and Cp == Cc and Kp == K and g.dim() == 4:
                return ops.conv2d_wgrad_tc_presplit(ops.split_f16_cached(x), ops.split_f16_cached(g), R, S, pad)
            dw = ops.conv2d_wgrad_tc(_pad_last(x, 8), _pad_last(g, 8), R, S, pad)
            return dw if (Cp == Cc and Kp == K) else dw[:K, :Cc].contiguous()
    if Cc <= WGRAD_TILE and K <= WGRAD_TILE:
        dw = ops.conv2d_wgrad(_pad_last(x, 8), _pad_last(g, 8), R, S, pad)
        return dw[:K, :Cc].contiguous()
    T = WGRAD_TILE
    xp, gp = _pad_last(x), _pad_last(g)
    xs = [(c0, min(T, Cc - c0), _pad_last(ops.chan_slice(xp, c0, min(T, xp.shape[-1] - c0)), 8))
          for c0 in range(0, Cc, T)] if Cc > T else [(0, Cc, _pad_last(x, 8))]
    gs = [(k0, min(T, K - k0), _pad_last(ops.chan_slice(gp, k0, min(T, gp.shape[-1] - k0)), 8))
          for k0 in range(0, K, T)] if K > T else [(0, K, _pad_last(g, 8))]
    dw = torch.empty(K, Cc, R, S, device=x.device)
    for c0, cn, xt in xs:
        for k0, kn, gt in gs:
            dw[k0:k0 + kn, c0:c0 + cn] = ops.conv2d_wgrad(xt, gt, R, S, pad)[:kn, :cn]
    return dw


def _flip_t(w):
    """[K,C,R,S] -> [C,K,R,S] rotated by 180 degrees: the weights of the data-gradient conv."""
    return w.flip(2, 3).transpose(0, 1)


TRAIN_TC_BIG_1X1 = os.environ.get("CRESTE_TRAIN_NO_TC_1X1") is None      # experiment / test switch
PRESPLIT_REUSE = True     # first-order training: split each conv operand once (forward x saved, g shared by dgrad / wgrad)


def _tc_f16(shape, K, R, S, pad):
    """True if a stride-1 conv of an [N,H,W,C] tensor runs in the 3xFP16 tensor-core mode with no channel padding."""
    # Large 1x1 convs with 64 <= C <= 192 AND >= 64 output channels over >= 64 K pixels (the dino head's 128 -> 128 /
    # 128 -> 256 layers at 16x128x240: 32 - 64 GFLOP each, 0.48 - 0.95 ms on the FFMA kernel, which the eval path keeps
    # for short reductions to save the operand pre-pass) run on the tensor cores in a training graph: the weight
    # gradient needs the split operands anyway.  Routing EVERY such conv there measured slower (profiles/r2d_experiments.md).
    big = TRAIN_TC_BIG_1X1 and K >= 64 and shape[0] * shape[1] * shape[2] >= 65536
    return (engine.get_precision() == "3xfp16" and shape[-1] % 8 == 0 and K % 8 == 0 and shape[1] * shape[2] >= 16
            and engine.pick_mode(tuple(shape), K, R, S, 1, pad, "3xfp16", prefer_tc=big) == "3xfp16")


def _conv_presplit(xs, w, pad):
    K, _, R, S = w.shape
    packed = ops.pack_conv_weight_f16_strided(w.detach().float())
    return ops.conv2d_presplit(xs, packed, K, R, S, 1, pad, precision="3xfp16")


class Conv2dFn(Function):
    @staticmethod
    def forward(ctx, x, w, ph, pw):
        ctx.pad = (ph, pw)
        K, Cc, R, S = w.shape
        pad = _pad4(ph, pw)
        # fast path (first-order training on the tensor cores): the operand is split once, used by this conv and SAVED
        # for the weight gradient; backward splits the output gradient once for the data and the weight gradient --
        # 2 operand pre-passes per conv instead of 4 (they were 11 % of a stage-1 step).  Same kernels, same scales:
        # bit-identical to the generic path.
        ctx.fast = bool(PRESPLIT_REUSE and x.dim() == 4 and _tc_f16(x.shape, K, R, S, pad)
                        and ops.wgrad_tc_supported(tuple(x.shape), K, R, S, pad))
        if ctx.fast:
            xs = ops.split_f16_cached(x)
            ctx.save_for_backward(x, w, xs.hi, xs.lo, xs.scal)
            return _conv_presplit(xs, w, pad)
        ctx.save_for_backward(x, w)
        return _conv_raw(x, w, ph, pw)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors[:2]
        ph, pw = ctx.pad
        R, S = w.shape[2], w.shape[3]
        dx = dw = None
        if ctx.fast and not torch.is_grad_enabled():      # no graph of the backward wanted (not a double backward)
            K, Cc = w.shape[0], w.shape[1]
            g = g.contiguous()
            gs = ops.split_f16(g)
            if ctx.needs_input_grad[0]:
                dpad = _pad4(_sub_pad(R - 1, ph), _sub_pad(S - 1, pw))
                if _tc_f16(g.shape, Cc, R, S, dpad):
                    dx = _conv_presplit(gs, _flip_t(w), dpad)
                else:
                    dx = _conv_raw(g, _flip_t(w), _sub_pad(R - 1, ph), _sub_pad(S - 1, pw))
            if ctx.needs_input_grad[1]:
                xs = ops.SplitAct(ctx.saved_tensors[2], ctx.saved_tensors[3], ctx.saved_tensors[4], x.shape)
                dw = ops.conv2d_wgrad_tc_presplit(xs, gs, R, S, _pad4(ph, pw))
            return dx, dw, None, None
        if ctx.needs_input_grad[0]:
            dx = Conv2dFn.apply(g, _flip_t(w), _sub_pad(R - 1, ph), _sub_pad(S - 1, pw))
        if ctx.needs_input_grad[1]:
            dw = WGradFn.apply(x, g, R, S, ph, pw)
        return dx, dw, None, None


def _sub_pad(k, p):
    """k - p for an int or a (low, high) padding pair (the data-gradient conv's padding)."""
    return k - p if isinstance(p, int) else (k - p[0], k - p[1])


class WGradFn(Function):
    @staticmethod
    def forward(ctx, x, g, R, S, ph, pw):
        ctx.save_for_backward(x, g)
        ctx.geom = (R, S, ph, pw)
        return _wgrad_raw(x, g, R, S, ph, pw)

    @staticmethod
    def backward(ctx, ggw):
        x, g = ctx.saved_tensors
        R, S, ph, pw = ctx.geom
        dx = dg = None
        if ctx.needs_input_grad[0]:
            dx = Conv2dFn.apply(g, _flip_t(ggw), R - 1 - ph, S - 1 - pw)
        if ctx.needs_input_grad[1]:
            dg = Conv2dFn.apply(x, ggw, ph, pw)
        return dx, dg, None, None, None, None


def _strided_dgrad(g, w, x_shape, stride, ph, pw):
    """Data gradient of a strided conv: stride-1 conv of the zero-inserted output gradient with the flipped
    weights (low pad R-1-pad, high pad chosen so that the result has the input's size)."""
    N, H, W, _ = x_shape
    _, P, Q, _ = g.shape
    R, S = w.shape[2], w.shape[3]
    Hz, Wz = stride * (P - 1) + 1, stride * (Q - 1) + 1
    z = ops.dilate(g, stride, Hz, Wz)
    return _conv_raw(z, _flip_t(w), (R - 1 - ph, H - Hz + ph), (S - 1 - pw, W - Wz + pw))


def _strided_wgrad(x, g, R, S, stride, ph, pw):
    """Weight gradient of a strided conv as stride-1 weight gradients over the stride^2 phase images:
    tap r with r - pad = stride * i + a (0 <= a < stride) is tap i of the phase image x[a::stride]."""
    N, H, W, Cc = x.shape
    _, P, Q, K = g.shape
    dw = torch.zeros(K, Cc, R, S, device=x.device, dtype=x.dtype)

    def taps(n, pad, a):
        rs = [r for r in range(n) if (r - pad) % stride == a]
        return rs, [(r - pad - a) // stride for r in rs]
    for a in range(stride):
        rs, iis = taps(R, ph, a)
        if not rs:
            continue
        for b in range(stride):
            ss, jjs = taps(S, pw, b)
            if not ss:
                continue
            Ri, Sj = iis[-1] - iis[0] + 1, jjs[-1] - jjs[0] + 1
            pt, pl = -iis[0], -jjs[0]
            xa = ops.phase_slice(x, stride, a, b, P + Ri - 1 - pt, Q + Sj - 1 - pl)
            dwa = _wgrad_raw(xa, g, Ri, Sj, (pt, 0), (pl, 0))               # [K,C,Ri,Sj]
            for r, i in zip(rs, iis):
                for s_, j in zip(ss, jjs):
                    dw[:, :, r, s_] = dwa[:, :, i - iis[0], j - jjs[0]]
    return dw


class StridedConvFn(Function):
    """Dense conv with stride > 1 (the ResNet-18 BEV trunk's 7x7 / 3x3 / 1x1 stride-2 layers, reference
    inpainting.py:80-90); first-order gradients only."""

    @staticmethod
    def forward(ctx, x, w, stride, ph, pw):
        ctx.save_for_backward(x, w)
        ctx.geom = (stride, ph, pw)
        return _conv_raw(x, w, ph, pw, stride)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        stride, ph, pw = ctx.geom
        g = g.contiguous()
        dx = _strided_dgrad(g, w, tuple(x.shape), stride, ph, pw) if ctx.needs_input_grad[0] else None
        dw = _strided_wgrad(x, g, w.shape[2], w.shape[3], stride, ph, pw) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None, None


class FrustumFn(Function):
    """depth [M,Hs,Ws] (m) -> voxel coordinates xy [M,P,2], height z [M,P], bounds mask [M,P] (uint8, no
    gradient): Camera2World + the bounds test + _points_to_voxels (splat_projection.py:19-51, :169, :175-189).
    xy and z are affine in the depth, so the backward is one elementwise kernel."""

    @staticmethod
    def forward(ctx, depth, p2p, rng, vox):
        xy, z, mask = ops.frustum_to_bev(depth, p2p, rng, vox)
        ctx.save_for_backward(p2p)
        ctx.geom = (tuple(depth.shape), vox)
        ctx.mark_non_differentiable(mask)
        return xy, z, mask

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gxy, gz, _gmask):
        (p2p,) = ctx.saved_tensors
        shape, vox = ctx.geom
        return ops.frustum_bwd(gxy, gz, p2p, shape, vox), None, None, None


class SplatFn(Function):
    """Bilinear BEV splat with mean normalisation (splat_projection.py:262-354) -> (bev NHWC [M,H,W,F],
    densities [M,1,H,W]); differentiable w.r.t. the point features and the voxel coordinates."""

    @staticmethod
    def forward(ctx, xy, feats, mask, H, W, min_weight):
        out = ops.splat_soft(xy, feats, mask, H, W, min_weight, want_nhwc=True, want_nchw=False)
        ctx.save_for_backward(xy, feats, mask, out["bev_nhwc"], out["dens"])
        ctx.min_weight = min_weight
        return out["bev_nhwc"], out["dens"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_bev, g_dens):
        xy, feats, mask, bev, dens = ctx.saved_tensors
        dfeats, dxy = ops.splat_soft_bwd(xy, feats, mask, bev, dens, g_bev, g_dens, ctx.min_weight)
        return dxy, dfeats.view_as(feats), None, None, None, None


class DepthExpectFn(Function):
    """Softmax expectation over the depth bins (depth_utils.py:300-313), logits NHWC [..., 128] -> [...]."""

    @staticmethod
    def forward(ctx, logits, dmin, dmax, out_div):
        ctx.save_for_backward(logits)
        ctx.cfg = (dmin, dmax, out_div)
        return ops.depth_expectation(logits, dmin, dmax, out_div)[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (logits,) = ctx.saved_tensors
        return ops.depth_expectation_bwd(logits, g, *ctx.cfg), None, None, None


def depth_expectation_nchw(logits_nchw, dmin, dmax, out_div=1000.0):
    return DepthExpectFn.apply(ToNHWC.apply(logits_nchw.float()), dmin, dmax, out_div)


class ChanAffineFn(Function):
    """y = act(x * a[c] + b[c]); a / b may be None (1 / 0)."""

    @staticmethod
    def forward(ctx, x, a, b, relu):
        y = ops.chan_affine(x, a, b, relu)
        ctx.relu = relu
        ctx.has_a = a is not None
        ctx.save_for_backward(x, a, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, a, y = ctx.saved_tensors
        gm = ReluBwdFn.apply(g, y) if ctx.relu else g
        dx = da = db = None
        if ctx.needs_input_grad[0]:
            dx = ChanAffineFn.apply(gm, a, None, False) if ctx.has_a else gm
        if ctx.has_a and ctx.needs_input_grad[1]:
            da = ChanDotFn.apply(gm, x)
        if ctx.needs_input_grad[2]:
            db = ChanDotFn.apply(gm, None)
        return dx, da, db, None


class ReluBwdFn(Function):
    """g * (y > 0); y carries no gradient."""

    @staticmethod
    def forward(ctx, g, y):
        ctx.save_for_backward(y)
        return ops.relu_bwd(g, y)

    @staticmethod
    def backward(ctx, gg):
        (y,) = ctx.saved_tensors
        return ReluBwdFn.apply(gg, y), None


class ChanDotFn(Function):
    """out[c] = sum_pix x[pix,c] * y[pix,c]   (y None: channel sum)."""

    @staticmethod
    def forward(ctx, x, y):
        ctx.save_for_backward(x, y)
        ctx.same = y is x
        return ops.chan_dot(x, y)

    @staticmethod
    def backward(ctx, gc):
        x, y = ctx.saved_tensors
        dx = dy = None
        if ctx.same:                      # sum x^2: one pass with 2*gc instead of two identical ones
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
                dx = ChanAffineFn.apply(x, 2.0 * gc, None, False)
            return dx, None
        if y is None:
            if ctx.needs_input_grad[0]:   # broadcast gc over the pixels
                zero = torch.zeros_like(gc)
                dx = ChanAffineFn.apply(x.detach(), zero, gc, False)
            return dx, None
        if ctx.needs_input_grad[0]:
            dx = ChanAffineFn.apply(y, gc, None, False)
        if ctx.needs_input_grad[1]:
            dy = ChanAffineFn.apply(x, gc, None, False)
        return dx, dy


class ChanStatsFn(Function):
    """(sum x, sum x^2) per channel in one pass, float64 [2, C].  Backward: d/dx = g1[c] + 2 g2[c] x,
    i.e. ONE ChanAffineFn -- so the Function set stays closed under differentiation."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.chan_stats(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ChanAffineFn.apply(x, (2.0 * g[1]).float(), g[0].float(), False)


class MaxPool2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.maxpool2(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return MaxPoolBwdFn.apply(g, x)


class MaxPoolBwdFn(Function):
    @staticmethod
    def forward(ctx, g, x):
        ctx.save_for_backward(x)
        return ops.maxpool2_bwd(x, g)

    @staticmethod
    def backward(ctx, gg):
        (x,) = ctx.saved_tensors
        return MaxPoolGatherFn.apply(gg, x), None


class MaxPoolGatherFn(Function):
    @staticmethod
    def forward(ctx, gg, x):
        ctx.save_for_backward(x)
        return ops.maxpool2_gather(x, gg)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return MaxPoolBwdFn.apply(g, x), None


class Up2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample2(x)

    @staticmethod
    def backward(ctx, g):
        return Up2AdjFn.apply(g)


class Up2AdjFn(Function):
    @staticmethod
    def forward(ctx, g):
        return ops.upsample2_adjoint(g)

    @staticmethod
    def backward(ctx, gg):
        return Up2Fn.apply(gg)


class ToNHWC(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.nchw_to_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return ToNCHW.apply(g)


class ToNCHW(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.nhwc_to_nchw(x)

    @staticmethod
    def backward(ctx, g):
        return ToNHWC.apply(g)


class RowDotFn(Function):
    """out[b] = sum_i x[b,i] * w[b,i] * mask[b,i]   (mask uint8 or None)."""

    @staticmethod
    def forward(ctx, x, w, mask):
        ctx.save_for_backward(x, w, mask)
        return ops.row_dot(x, w, mask)

    @staticmethod
    def backward(ctx, gb):
        x, w, mask = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = RowScaleFn.apply(w, gb, mask)
        if ctx.needs_input_grad[1]:
            dw = RowScaleFn.apply(x, gb, mask)
        return dx, dw, None


class RowScaleFn(Function):
    """out[b,i] = x[b,i] * s[b] * mask[b,i]."""

    @staticmethod
    def forward(ctx, x, s, mask):
        ctx.save_for_backward(x, s, mask)
        return ops.row_scale(x, s, mask)

    @staticmethod
    def backward(ctx, gg):
        x, s, mask = ctx.saved_tensors
        dx = ds = None
        if ctx.needs_input_grad[0]:
            dx = RowScaleFn.apply(gg, s, mask)
        if ctx.needs_input_grad[1]:
            ds = RowDotFn.apply(gg, x, mask)
        return dx, ds, None


class GradPenaltyFn(Function):
    """mean_{b,pixel} (||G[b,:,pixel]||_2 - 1)^2 for G NCHW (loss_utils.py:1216-1217)."""

    @staticmethod
    def forward(ctx, G):
        ctx.save_for_backward(G)
        return ops.grad_penalty(G)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (G,) = ctx.saved_tensors
        return ops.grad_penalty_bwd(G, g)


# ----------------------------------------------------------------------------- composed layers
def conv2d(x, conv):
    """nn.Conv2d parameters -> differentiable stride-1 NHWC conv (+ bias)."""
    stride = conv.stride[0] if isinstance(conv.stride, tuple) else conv.stride
    ph, pw = conv.padding if isinstance(conv.padding, tuple) else (conv.padding,) * 2
    if stride != 1:
        y = StridedConvFn.apply(x, conv.weight, int(stride), int(ph), int(pw))     # first-order only
    else:
        y = Conv2dFn.apply(x, conv.weight, int(ph), int(pw))
    if conv.bias is not None:
        y = ChanAffineFn.apply(y, None, conv.bias, False)
    return y


def relu(x):
    return ChanAffineFn.apply(x, None, None, True)


def batch_norm(x, bn, relu=False):
    """nn.BatchNorm2d over channels-last x, honouring bn.training exactly like F.batch_norm:
    batch statistics (biased variance) + running-stat update in training mode, running statistics
    in eval mode.  Composed of ChanDotFn / ChanAffineFn so it is differentiable to any order."""
    Cc = x.shape[-1]
    M = x.numel() // Cc
    use_batch = bn.training or bn.running_mean is None
    if use_batch:
        # one pass for both moments (double accumulators, so E[x^2] - mean^2 is safe), then ONE
        # affine pass: y = x * (gamma * inv) + (beta - mean * gamma * inv).  The [C]-sized algebra in
        # between is float64 torch arithmetic on <= 64 numbers.
        st = ChanStatsFn.apply(x)
        mean = st[0] / M
        var = (st[1] / M - mean * mean).clamp_min(0.0)
        if bn.training and bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - mom).add_(mean.detach().float(), alpha=mom)
                bn.running_var.mul_(1 - mom).add_((var.detach() * (M / max(M - 1, 1))).float(), alpha=mom)
        inv = torch.rsqrt(var + bn.eps)
        scale = inv * bn.weight.double() if bn.affine else inv
        shift = (bn.bias.double() if bn.affine else 0.0) - mean * scale
        return ChanAffineFn.apply(x, scale.float(), shift.float(), relu)
    inv = torch.rsqrt(bn.running_var + bn.eps)
    scale = inv * bn.weight if bn.affine else inv
    shift = (bn.bias if bn.affine else 0) - bn.running_mean * scale
    return ChanAffineFn.apply(x, scale, shift, relu)


# ===================================================================== stage-1 backbone training
# Train-mode graph of DistillationBackbone (reference creste/models/distillation.py:145-207 driven by
# creste/train_pefree.py:76-106).  The reference needs first-order gradients only here, so the
# fused nodes below are `once_differentiable`; dense convs reuse Conv2dFn / WGradFn above.
once = torch.autograd.function.once_differentiable


def _update_running(bn, mean, var, M):
    """F.batch_norm(training=True) bookkeeping: momentum update with the UNBIASED variance."""
    if not (bn.track_running_stats and bn.running_mean is not None):
        return
    with torch.no_grad():
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        bn.running_mean.mul_(1 - mom).add_(mean.float(), alpha=mom)
        bn.running_var.mul_(1 - mom).add_((var * (M / max(M - 1, 1))).float(), alpha=mom)


FUSED_BN_BWD = True       # BatchNorm(+act) backward: second pass recomputes gu instead of reading it back


class BNActFn(Function):
    """BatchNorm2d with batch statistics + activation ('none' | 'relu' | 'swish') as ONE node:
    forward = one moments pass + one [C]-sized finalize launch + one affine/act pass; backward = one pass
    producing gu = g*act'(u) with (sum gu, sum gu*x), one finalize launch, then dx = gu*a + x*q + r."""

    @staticmethod
    def forward(ctx, x, weight, bias, bn, act):
        Cc = x.shape[-1]
        M = x.numel() // Cc
        st = ops.chan_moments(x)
        track = bn.training and bn.track_running_stats and bn.running_mean is not None
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        mom = bn.momentum
        if track and mom is None:
            mom = 1.0 / float(bn.num_batches_tracked)        # cumulative average (one host sync, as torch)
        ab, mi = ops.bn_fwd_finalize(st, weight.detach() if weight is not None else None,
                                     bias.detach() if bias is not None else None, M, bn.eps, mom,
                                     bn.running_mean if track else None, bn.running_var if track else None)
        if track:       # updated in place through raw pointers by the finalize kernel
            engine.mark_written([bn.running_mean, bn.running_var])
        ctx.save_for_backward(x, ab, mi)
        ctx.act, ctx.M = act, M
        # 3xFP16 training: the affine / activation pass also measures max|y|, so the conv that consumes y (and the
        # data- / weight-gradient convs that consume dx below) split their operand without an amax pass
        return ops.chan_affine_act(x, ab[0], ab[1], act, want_amax=(engine.get_precision() == "3xfp16"))

    @staticmethod
    @once
    def backward(ctx, g):
        x, ab, mi = ctx.saved_tensors
        # gu = g * act'(u) is not stored: the second pass recomputes it from g (20 instead of 24 bytes per element)
        fused = ctx.act != "none" and FUSED_BN_BWD
        gu, sums = ops.bn_act_bwd(g, x, ab[0], ab[1], ctx.act, want_gu=not fused)
        out4 = ops.bn_bwd_finalize(sums, ab, mi, ctx.M)
        pub = engine.get_precision() == "3xfp16"
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (ops.chan_axpby_act(g, x, ab[0], ab[1], ctx.act, ab[0], out4[2], out4[3], want_amax=pub) if fused
                  else ops.chan_axpby(gu, x, ab[0], out4[2], out4[3], want_amax=pub))
        return dx, out4[0], out4[1], None, None


def bn_act(x, bn, act="none"):
    """nn.BatchNorm2d (+ activation) over channels-last x inside the stage-1 training graph."""
    if bn.training or bn.running_mean is None:
        return BNActFn.apply(x, bn.weight, bn.bias, bn, act)
    if act == "swish":
        raise NotImplementedError("eval-mode BatchNorm + swish inside a training graph (frozen trunk) is unused")
    return batch_norm(x, bn, relu=(act == "relu"))


class DwConvFn(Function):
    """Depthwise k x k conv, weights [C,1,k,k] (torch layout), pad = (top, bottom, left, right)."""

    @staticmethod
    def forward(ctx, x, w, k, stride, pad):
        w_rsc = w.detach().permute(2, 3, 1, 0).reshape(k * k, w.shape[0]).contiguous()
        ctx.save_for_backward(x, w_rsc)
        ctx.geom = (k, stride, pad)
        return ops.dwconv_fwd(x, w_rsc, k, stride, pad)

    @staticmethod
    @once
    def backward(ctx, g):
        x, w_rsc = ctx.saved_tensors
        k, stride, pad = ctx.geom
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = ops.dwconv_dgrad(g, w_rsc, tuple(x.shape), k, stride, pad)
        if ctx.needs_input_grad[1]:
            dw = ops.dwconv_wgrad(x, g, k, stride, pad).view(k, k, 1, -1).permute(3, 2, 0, 1).contiguous()
        return dx, dw, None, None, None


class StemConvFn(Function):
    """The strided C=4 stem conv (efficientnet `_conv_stem`): exact-fp32 forward, weight gradient
    only (its input is the image)."""

    @staticmethod
    def forward(ctx, x, w, stride, pad):
        K, Cc, R, S = w.shape
        ctx.save_for_backward(x)
        ctx.geom = (R, S, stride, pad)
        return ops.conv2d(x, ops.pack_conv_weight(w.detach().float()), K, R, S, stride, pad, precision="fp32")

    @staticmethod
    @once
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        R, S, stride, pad = ctx.geom
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("StemConvFn: the data gradient of the strided stem is never needed")
        return None, ops.wgrad_strided(x, g, R, S, stride, pad), None, None


class SamplePoolFn(Function):
    """adaptive_avg_pool2d(x, 1) over NHWC -> [B,1,1,C]."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        HW = x.numel() // (x.shape[0] * x.shape[-1])
        return ops.sample_dot(x, None, 1.0 / HW)

    @staticmethod
    @once
    def backward(ctx, g):
        shape = ctx.shape
        HW = 1
        for d in shape[1:-1]:
            HW *= d
        return ops.sample_affine(None, None, (g / HW).contiguous(), shape=shape)


class SampleScaleFn(Function):
    """x * gate[b,c] (squeeze-excite)."""

    @staticmethod
    def forward(ctx, x, gate):
        ctx.save_for_backward(x, gate)
        return ops.sample_affine(x, gate)

    @staticmethod
    @once
    def backward(ctx, g):
        x, gate = ctx.saved_tensors
        dx = ops.sample_affine(g, gate) if ctx.needs_input_grad[0] else None
        dgate = ops.sample_dot(g, x).view_as(gate) if ctx.needs_input_grad[1] else None
        return dx, dgate


class ActFn(Function):
    """swish / sigmoid on the small squeeze-excite tensors."""

    @staticmethod
    def forward(ctx, x, kind):
        ctx.save_for_backward(x)
        ctx.kind = kind
        return ops.act(x, kind)

    @staticmethod
    @once
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.act_bwd(g, x, ctx.kind), None


class AddScaledFn(Function):
    """inp + x * s[b]: identity skip with drop-connect (s None: plain residual sum)."""

    @staticmethod
    def forward(ctx, x, inp, s):
        ctx.save_for_backward(s)
        return ops.add_scaled(inp, x, s)

    @staticmethod
    @once
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        dx = g if s is None else ops.row_scale(g, s)
        return dx, g, None


class UpCatFn(Function):
    """cat([skip, bilinear_x2(x)], C) (reference effnet.py:22-25) and its adjoint."""

    @staticmethod
    def forward(ctx, x, skip, sf):
        if isinstance(sf, (tuple, list)) or int(sf) != sf:
            raise NotImplementedError("training path: integer up-sampling factors only (2 in the RGB-D decoder, "
                                      "4 in the BEV DeconvHeads)")
        sf = int(sf)
        ctx.cs, ctx.cx, ctx.sf = skip.shape[-1], x.shape[-1], sf
        ctx.hw = (x.shape[1], x.shape[2])
        return ops.upsample_concat(skip, x, (sf * x.shape[1], sf * x.shape[2]), sf)

    @staticmethod
    @once
    def backward(ctx, g):
        dskip = ops.chan_slice(g, 0, ctx.cs) if ctx.needs_input_grad[1] else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.upsample_adjoint_slice(g, ctx.cs, ctx.cx, ctx.hw[0], ctx.hw[1], 1.0 / ctx.sf)
        return dx, dskip, None


class CEDepthFn(Function):
    """mean cross-entropy over the valid pixels (CrossEntropyDepth, loss_utils.py:477-527); `acc` is
    the device float64[>=2] {sum CE, #valid} from creste_stage1_depth_losses."""

    @staticmethod
    def forward(ctx, logits_nchw, label_mm, acc, dmin, dmax):
        ctx.save_for_backward(logits_nchw, label_mm, acc)
        ctx.rng = (dmin, dmax)
        return (acc[0] / acc[1]).float()

    @staticmethod
    @once
    def backward(ctx, g):
        logits, label, acc = ctx.saved_tensors
        scale = (g.double() / acc[1]).float()
        return ops.ce_depth_bwd(logits, label, ctx.rng[0], ctx.rng[1], scale), None, None, None, None


class MaskedMSEFn(Function):
    """mean((pred - gt)^2) over the elements whose target is finite (MSELoss, loss_utils.py:606-647)."""

    @staticmethod
    def forward(ctx, pred, gt):
        acc = ops.masked_mse(pred, gt)
        ctx.save_for_backward(pred, gt, acc)
        return (acc[0] / acc[1]).float()

    @staticmethod
    @once
    def backward(ctx, g):
        pred, gt, acc = ctx.saved_tensors
        scale = (2.0 * g.double() / acc[1]).float()
        return ops.masked_mse_bwd(pred, gt, scale), None


# ============================================================================ stage-2 (train_ssc.py) losses
class SmoothL1Fn(Function):
    """mean Smooth-L1(pred - gt * gt_scale) over mask & isfinite(gt) (loss_utils.py:530-604)."""

    @staticmethod
    def forward(ctx, pred, gt, mask, gt_scale, beta):
        acc = ops.smooth_l1(pred, gt, mask, gt_scale, beta)
        ctx.save_for_backward(pred, gt, mask, acc)
        ctx.cfg = (gt_scale, beta)
        return (acc[0] / acc[1]).float()

    @staticmethod
    @once
    def backward(ctx, g):
        pred, gt, mask, acc = ctx.saved_tensors
        scale = (g.double() / acc[1]).float()
        return ops.smooth_l1_bwd(pred, gt, mask, ctx.cfg[0], ctx.cfg[1], scale).view_as(pred), None, None, None, None


class WeightedCEFn(Function):
    """torch.nn.CrossEntropyLoss(weight, ignore_index, reduction='mean') over the masked cells of NCHW logits
    (loss_utils.py:379-474); `acc` is the float64[4] result of creste_ce_weighted."""

    @staticmethod
    def forward(ctx, logits, labels, mask, weights, ignore_index, acc):
        ctx.save_for_backward(logits, labels, mask, weights, acc)
        ctx.ignore_index = ignore_index
        return (acc[0] / acc[1]).float()

    @staticmethod
    @once
    def backward(ctx, g):
        logits, labels, mask, weights, acc = ctx.saved_tensors
        scale = (g.double() / acc[1]).float()
        return ops.ce_weighted_bwd(logits, labels, mask, weights, ctx.ignore_index, scale), None, None, None, None, None


class L2NormRowsFn(Function):
    @staticmethod
    def forward(ctx, x):
        y, nrm = ops.l2norm_rows(x)
        ctx.save_for_backward(y, nrm)
        return y

    @staticmethod
    @once
    def backward(ctx, g):
        y, nrm = ctx.saved_tensors
        return ops.l2norm_rows_bwd(y, g, nrm)


class SupConFn(Function):
    """Multi-positive contrastive loss (supcon_loss.py:56-115) of local rows `f` against the gathered rows `a`:
    the N x Na similarity matrix is never materialised.  Returns the mean over the local rows; backward gives
    d f (rows) and d a (columns) -- the caller's differentiable all-gather turns d a into a reduce-scatter."""

    @staticmethod
    def forward(ctx, f, a, lf, la, self_off, temperature, class_weights):
        stats, acc = ops.supcon_fwd(f, a, lf, la, self_off, temperature, class_weights)
        ctx.save_for_backward(f, a, lf, la, class_weights, stats)
        ctx.cfg = (self_off, temperature)
        return (acc[0] / f.shape[0]).float()

    @staticmethod
    @once
    def backward(ctx, g):
        f, a, lf, la, cw, stats = ctx.saved_tensors
        scale = (g / f.shape[0]).float()
        df, da = ops.supcon_bwd(f, a, lf, la, ctx.cfg[0], ctx.cfg[1], cw, stats, scale)
        return df, da, None, None, None, None, None
