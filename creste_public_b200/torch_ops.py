"""`torch.library` registration of the eval-forward kernels: the export path.

The reference's only deployment interface is scripts/runtime/compile.py:160-210 --
`torch.jit.trace(model, (inputs,), strict=False)` -> `.save()` -> loaded by the C++ runtime.  A ctypes call is
invisible to the tracer, so every C-ABI entry point the eval forward uses is ALSO a dispatcher op in the `creste::`
namespace: under `torch.jit.trace` the wrappers of creste_public_b200.ops route through `torch.ops.creste.*`, the
trace records those calls (packed weights / folded BatchNorm factors become graph constants), and the saved module
runs wherever the `creste::` ops are registered (importing this module; a C++ loader registers the same schemas over
libcreste_b200.so).  Outside tracing the wrappers call the C ABI directly: no dispatcher overhead on the hot path.

Schemas (all tensors CUDA fp32 unless noted; NHWC activations):
  creste::conv2d(x, w, K, R, S, stride, pad[4], scale?, shift?, gate?, residual?, act, out_nchw, precision) -> Tensor
  creste::dwconv_bn_swish(x, w, scale, shift, R, stride, pad[4]) -> (out, chan_part)
  creste::se_gate(chan_part, hw, w_red, b_red, w_exp, b_exp) -> gate
  creste::upsample_concat(skip?, x, out_hw[2], ratio_h, ratio_w, x_first) -> Tensor
  creste::maxpool2_concat(srcs[], rows_out) -> (nhwc, nchw)
  creste::nchw_to_nhwc(x) / creste::nhwc_to_nchw(x) -> Tensor
  creste::depth_expectation(logits, dmin, dmax, out_div) -> (metric, bins int64)
  creste::frustum_to_bev(depth, p2p, range[6], voxel[2]) -> (xy, z, mask uint8)
  creste::zmlp_concat(feats, z, w1, b1, w2, b2) -> Tensor
  creste::splat_soft(xy, feats, mask?, H, W, min_weight) -> (bev_nhwc, bev_nchw, dens)
  creste::proj_head(x, w, bias?) -> (pred_nhwc, pred_nchw, x_nchw)
"""
import threading
from typing import List, Optional, Tuple

import torch
from torch import Tensor

_tls = threading.local()


def inside():
    """True while a creste:: op implementation is executing (its own ops.* calls must not re-dispatch)."""
    return getattr(_tls, "depth", 0) > 0


class _Enter:
    def __enter__(self):
        _tls.depth = getattr(_tls, "depth", 0) + 1

    def __exit__(self, *a):
        _tls.depth -= 1


def _raw():
    from . import ops
    return ops._RAW


@torch.library.custom_op("creste::conv2d", mutates_args=())
def conv2d(x: Tensor, w: Tensor, K: int, R: int, S: int, stride: int, pad: List[int], scale: Optional[Tensor],
           shift: Optional[Tensor], gate: Optional[Tensor], residual: Optional[Tensor], act: str, out_nchw: bool,
           precision: str) -> Tensor:
    with _Enter():
        return _raw()["conv2d"](x, w, K, R, S, stride, tuple(pad), scale, shift, gate, residual, act, out_nchw, precision)


@conv2d.register_fake
def _(x, w, K, R, S, stride, pad, scale, shift, gate, residual, act, out_nchw, precision):
    N, H, W, _ = x.shape
    P = (H + pad[0] + pad[1] - R) // stride + 1
    Q = (W + pad[2] + pad[3] - S) // stride + 1
    return x.new_empty((N, K, P, Q) if out_nchw else (N, P, Q, K))


@torch.library.custom_op("creste::dwconv_bn_swish", mutates_args=())
def dwconv_bn_swish(x: Tensor, w: Tensor, scale: Tensor, shift: Tensor, R: int, stride: int,
                    pad: List[int]) -> Tuple[Tensor, Tensor]:
    with _Enter():
        return _raw()["dwconv_bn_swish"](x, w, scale, shift, R, stride, tuple(pad))


@dwconv_bn_swish.register_fake
def _(x, w, scale, shift, R, stride, pad):
    N, H, W, C = x.shape
    P = (H + pad[0] + pad[1] - R) // stride + 1
    Q = (W + pad[2] + pad[3] - R) // stride + 1
    return x.new_empty(N, P, Q, C), x.new_empty(N, 1, C)


@torch.library.custom_op("creste::se_gate", mutates_args=())
def se_gate(chan_part: Tensor, hw: int, w_red: Tensor, b_red: Tensor, w_exp: Tensor, b_exp: Tensor) -> Tensor:
    with _Enter():
        return _raw()["se_gate"](chan_part, hw, w_red, b_red, w_exp, b_exp)


@se_gate.register_fake
def _(chan_part, hw, w_red, b_red, w_exp, b_exp):
    return chan_part.new_empty(chan_part.shape[0], chan_part.shape[2])


@torch.library.custom_op("creste::upsample_concat", mutates_args=())
def upsample_concat(skip: Optional[Tensor], x: Tensor, out_hw: List[int], ratio_h: float, ratio_w: float,
                    x_first: bool) -> Tensor:
    with _Enter():
        return _raw()["upsample_concat"](skip, x, tuple(out_hw), (1.0 / ratio_h, 1.0 / ratio_w), x_first)


@upsample_concat.register_fake
def _(skip, x, out_hw, ratio_h, ratio_w, x_first):
    Cs = 0 if skip is None else skip.shape[-1]
    return x.new_empty(x.shape[0], out_hw[0], out_hw[1], Cs + x.shape[-1])


@torch.library.custom_op("creste::maxpool2_concat", mutates_args=())
def maxpool2_concat(srcs: List[Tensor], rows_out: int) -> Tuple[Tensor, Tensor]:
    with _Enter():
        return _raw()["maxpool2_concat"](list(srcs), rows_out, True)


@maxpool2_concat.register_fake
def _(srcs, rows_out):
    N, H, W, _ = srcs[0].shape
    Ct = sum(int(s.shape[-1]) for s in srcs)
    return srcs[0].new_empty(N, rows_out, W // 2, Ct), srcs[0].new_empty(N, Ct, rows_out, W // 2)


@torch.library.custom_op("creste::nchw_to_nhwc", mutates_args=())
def nchw_to_nhwc(x: Tensor) -> Tensor:
    with _Enter():
        return _raw()["nchw_to_nhwc"](x)


@nchw_to_nhwc.register_fake
def _(x):
    N, C, H, W = x.shape
    return x.new_empty(N, H, W, C)


@torch.library.custom_op("creste::nhwc_to_nchw", mutates_args=())
def nhwc_to_nchw(x: Tensor) -> Tensor:
    with _Enter():
        return _raw()["nhwc_to_nchw"](x)


@nhwc_to_nchw.register_fake
def _(x):
    N, H, W, C = x.shape
    return x.new_empty(N, C, H, W)


@torch.library.custom_op("creste::depth_expectation", mutates_args=())
def depth_expectation(logits: Tensor, dmin: float, dmax: float, out_div: float) -> Tuple[Tensor, Tensor]:
    with _Enter():
        return _raw()["depth_expectation"](logits, dmin, dmax, out_div)


@depth_expectation.register_fake
def _(logits, dmin, dmax, out_div):
    return logits.new_empty(logits.shape[:-1]), logits.new_empty(logits.shape[:-1], dtype=torch.int64)


@torch.library.custom_op("creste::frustum_to_bev", mutates_args=())
def frustum_to_bev(depth: Tensor, p2p: Tensor, pc_range: List[float], voxel: List[float]) -> Tuple[Tensor, Tensor, Tensor]:
    with _Enter():
        return _raw()["frustum_to_bev"](depth, p2p, list(pc_range), list(voxel))


@frustum_to_bev.register_fake
def _(depth, p2p, pc_range, voxel):
    N, Hs, Ws = depth.shape
    return (depth.new_empty(N, Hs * Ws, 2), depth.new_empty(N, Hs * Ws),
            depth.new_empty(N, Hs * Ws, dtype=torch.uint8))


@torch.library.custom_op("creste::zmlp_concat", mutates_args=())
def zmlp_concat(feats: Tensor, z: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor) -> Tensor:
    with _Enter():
        return _raw()["zmlp_concat"](feats, z, w1, b1, w2, b2)


@zmlp_concat.register_fake
def _(feats, z, w1, b1, w2, b2):
    return feats.new_empty(*feats.shape[:-1], feats.shape[-1] + 32)


@torch.library.custom_op("creste::splat_soft", mutates_args=())
def splat_soft(xy: Tensor, feats: Tensor, mask: Optional[Tensor], H: int, W: int,
               min_weight: float) -> Tuple[Tensor, Tensor, Tensor]:
    with _Enter():
        o = _raw()["splat_soft"](xy, feats, mask, H, W, min_weight, True, True, False)
    return o["bev_nhwc"], o["bev_nchw"], o["dens"]


@splat_soft.register_fake
def _(xy, feats, mask, H, W, min_weight):
    N, F = xy.shape[0], feats.shape[-1]
    return xy.new_empty(N, H, W, F), xy.new_empty(N, F, H, W), xy.new_empty(N, 1, H, W)


@torch.library.custom_op("creste::proj_head", mutates_args=())
def proj_head(x: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
    with _Enter():
        return _raw()["proj_head"](x, w, bias, True, True)


@proj_head.register_fake
def _(x, w, bias):
    N, H, W, C = x.shape
    K = w.shape[0]
    return x.new_empty(N, H, W, K), x.new_empty(N, K, H, W), x.new_empty(N, C, H, W)


NAMES = ["conv2d", "dwconv_bn_swish", "se_gate", "upsample_concat", "maxpool2_concat", "nchw_to_nhwc", "nhwc_to_nchw",
         "depth_expectation", "frustum_to_bev", "zmlp_concat", "splat_soft", "proj_head"]
