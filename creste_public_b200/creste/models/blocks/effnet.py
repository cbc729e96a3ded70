"""EfficientNet-B0 RGB-D encoder + U-Net decoder: mirror of reference
creste/models/blocks/effnet.py (`Up` :8-28, `EffNet` :31-98) and of the EfficientNet-B0 trunk the
reference takes from the third-party `efficientnet_pytorch` package (not vendored, not pinned:
SURVEY.md section 8(c)); `EfficientNetB0` below carries that library's parameter names
(`_conv_stem`, `_bn0`, `_blocks.{i}._expand_conv`, ... `_conv_head`, `_bn1`, `_fc`) so reference
checkpoints load unchanged.

Compute (eval mode): NHWC end to end.
  1x1 expand / project, stem, decoder 3x3 convs  -> creste_conv2d (BN folded, swish / ReLU fused,
                                                    SE gate folded into the project conv's A load,
                                                    identity skip fused as the residual)
  depthwise 3x3 / 5x5 + BN + swish + SE pooling  -> creste_dwconv_bn_swish (one pass)
  SE squeeze / excite                            -> creste_se_gate
  bilinear x2 + concat                           -> creste_upsample_concat
The dead `_conv_head` (320->1280, never read by EffNet.forward, reference effnet.py:85-88) is
kept as parameters but never executed.
"""
import math

import torch
from torch import nn

from creste_public_b200 import autograd as ag
from creste_public_b200 import engine, ops
from creste_public_b200.engine import FusedConv, PackCache, bn_scale_shift, require_eval

# (repeats, kernel, stride, expand, in, out) -- EfficientNet-B0, SE ratio 0.25
B0_STAGES = [(1, 3, 1, 1, 32, 16), (2, 3, 2, 6, 16, 24), (2, 5, 2, 6, 24, 40), (3, 3, 2, 6, 40, 80),
             (3, 5, 1, 6, 80, 112), (4, 5, 2, 6, 112, 192), (1, 3, 1, 6, 192, 320)]
BN_MOM, BN_EPS = 0.01, 1e-3


def same_pad(k, s):
    """Static TF-'SAME' padding as efficientnet_pytorch fixes it at construction (its nominal
    image sizes are even at every strided layer): (low, high) per spatial dim."""
    total = (k - 1) if s == 1 else max(k - s, 0)
    return total // 2, total - total // 2


class MBConv(nn.Module):
    def __init__(self, k, s, e, cin, cout):
        super().__init__()
        self.k, self.s, self.e, self.cin, self.cout = k, s, e, cin, cout
        mid = cin * e
        if e != 1:
            self._expand_conv = nn.Conv2d(cin, mid, 1, bias=False)
            self._bn0 = nn.BatchNorm2d(mid, momentum=BN_MOM, eps=BN_EPS)
        self._depthwise_conv = nn.Conv2d(mid, mid, k, stride=s, groups=mid, bias=False)
        self._bn1 = nn.BatchNorm2d(mid, momentum=BN_MOM, eps=BN_EPS)
        nsq = max(1, int(cin * 0.25))
        self._se_reduce = nn.Conv2d(mid, nsq, 1)
        self._se_expand = nn.Conv2d(nsq, mid, 1)
        self._project_conv = nn.Conv2d(mid, cout, 1, bias=False)
        self._bn2 = nn.BatchNorm2d(cout, momentum=BN_MOM, eps=BN_EPS)
        self._cache = PackCache()
        if e != 1:
            object.__setattr__(self, "_f_expand", FusedConv(self._expand_conv, self._bn0))
        object.__setattr__(self, "_f_project", FusedConv(self._project_conv, self._bn2))

    def forward_train(self, x, drop_rate):
        """Train-mode MBConv as an autograd graph of sm_100a kernels (BatchNorm batch statistics,
        squeeze-excite, drop-connect on the identity skip): efficientnet_pytorch MBConvBlock.forward."""
        inp = x
        if self.e != 1:
            x = ag.bn_act(ag.conv2d(x, self._expand_conv), self._bn0, "swish")
        lo, hi = same_pad(self.k, self.s)
        x = ag.DwConvFn.apply(x, self._depthwise_conv.weight, self.k, self.s, (lo, hi, lo, hi))
        x = ag.bn_act(x, self._bn1, "swish")
        sq = ag.SamplePoolFn.apply(x)
        sq = ag.ActFn.apply(ag.conv2d(sq, self._se_reduce), "swish")
        gate = ag.ActFn.apply(ag.conv2d(sq, self._se_expand), "sigmoid")
        x = ag.SampleScaleFn.apply(x, gate)
        x = ag.bn_act(ag.conv2d(x, self._project_conv), self._bn2, "none")
        if self.s == 1 and self.cin == self.cout:
            scale = None
            if drop_rate and self.training:
                scale = engine.drop_connect_scale(x.shape[0], drop_rate, x.device)
            x = ag.AddScaledFn.apply(x, inp, scale)
        return x

    def forward_nhwc(self, x):
        inp = x
        if self.e != 1:
            x = self._f_expand(x, act="swish")
        dw, bn = self._depthwise_conv, self._bn1
        mid = dw.weight.shape[0]

        def build_dw():
            w = dw.weight.detach().float().permute(2, 3, 1, 0).reshape(self.k * self.k, mid).contiguous()
            return (w,) + bn_scale_shift(bn)
        w, scale, shift = self._cache.get("dw", [dw.weight, bn.weight, bn.bias, bn.running_mean,
                                                 bn.running_var], build_dw)
        lo, hi = same_pad(self.k, self.s)
        # max|dw out| travels with the tensor when the project conv runs on the fp16 tensor-core path: it bounds
        # max|out * gate| (sigmoid gate), so the operand pre-pass of that conv makes no amax pass
        track = engine.get_precision() in ("3xfp16", "fp16") and not torch.jit.is_tracing()
        amax = torch.empty(1, device=x.device) if track else None
        x, csum = ops.dwconv_bn_swish(x, w, scale, shift, self.k, self.s, (lo, hi, lo, hi),
                                      **({"amax_out": amax} if track else {}))
        if track:
            x._amax = amax

        def build_se():
            r, e = self._se_reduce, self._se_expand
            return (r.weight.detach().float().reshape(r.weight.shape[0], mid).contiguous(),
                    r.bias.detach().float().contiguous(),
                    e.weight.detach().float().reshape(mid, e.weight.shape[1]).contiguous(),
                    e.bias.detach().float().contiguous())
        wr, br, we, be = self._cache.get("se", [self._se_reduce.weight, self._se_reduce.bias,
                                                self._se_expand.weight, self._se_expand.bias], build_se)
        gate = ops.se_gate(csum, x.shape[1] * x.shape[2], wr, br, we, be)
        skip = inp if (self.s == 1 and self.cin == self.cout) else None
        return self._f_project(x, act="none", gate=gate, residual=skip)


class EfficientNetB0(nn.Module):
    """Parameter tree of efficientnet_pytorch's EfficientNet('efficientnet-b0')."""

    def __init__(self, in_channels=3):
        super().__init__()
        self._conv_stem = nn.Conv2d(in_channels, 32, 3, stride=2, bias=False)
        self._bn0 = nn.BatchNorm2d(32, momentum=BN_MOM, eps=BN_EPS)
        blocks = []
        for (rep, k, s, e, cin, cout) in B0_STAGES:
            for r in range(rep):
                blocks.append(MBConv(k, s if r == 0 else 1, e, cin if r == 0 else cout, cout))
        self._blocks = nn.ModuleList(blocks)
        self._conv_head = nn.Conv2d(320, 1280, 1, bias=False)       # dead in EffNet.forward
        self._bn1 = nn.BatchNorm2d(1280, momentum=BN_MOM, eps=BN_EPS)
        self._fc = nn.Linear(1280, 1000)                             # dead, kept for checkpoints

    def set_swish(self, memory_efficient=True):
        return None

    def _stem(self):
        if "_f_stem" not in self.__dict__:
            object.__setattr__(self, "_f_stem", FusedConv(self._conv_stem, self._bn0))
        return self.__dict__["_f_stem"]

    DROP_CONNECT = 0.2      # efficientnet-b0 global drop_connect_rate

    def extract_endpoints_train(self, x):
        """extract_endpoints in training mode: drop-connect rate 0.2 * idx / n per block, and the
        1280-channel head runs (without gradient) only because the reference's BatchNorm `_bn1`
        updates its running statistics from it."""
        lo, hi = same_pad(3, 2)
        x = ag.StemConvFn.apply(x, self._conv_stem.weight, 2, (lo, hi, lo, hi))
        x = ag.bn_act(x, self._bn0, "swish")
        endpoints = {}
        prev = x
        n = len(self._blocks)
        for idx, blk in enumerate(self._blocks):
            x = blk.forward_train(x, self.DROP_CONNECT * float(idx) / n)
            if prev.shape[1] > x.shape[1]:
                endpoints[f"reduction_{len(endpoints) + 1}"] = prev
            elif idx == n - 1:
                endpoints[f"reduction_{len(endpoints) + 1}"] = x
            prev = x
        if self._bn1.training:
            with torch.no_grad():
                h = ag.conv2d(x.detach(), self._conv_head)
                M = h.numel() // h.shape[-1]
                st = ops.chan_moments(h)
                mean = st[0] / M
                ag._update_running(self._bn1, mean, (st[1] / M - mean * mean).clamp_min(0.0), M)
        return endpoints

    def extract_endpoints_nhwc(self, x):
        """Endpoint rule of efficientnet_pytorch.extract_endpoints: the activation *before* every
        resolution drop, plus the last block's output (reduction_1..5); the 1280-ch head
        (reduction_6) is not computed because nothing reads it."""
        lo, hi = same_pad(3, 2)
        x = self._stem()(x, act="swish", pad=(lo, hi, lo, hi))
        endpoints = {}
        prev = x
        n = len(self._blocks)
        for idx, blk in enumerate(self._blocks):
            x = blk.forward_nhwc(x)
            if prev.shape[1] > x.shape[1]:
                endpoints[f"reduction_{len(endpoints) + 1}"] = prev
            elif idx == n - 1:
                endpoints[f"reduction_{len(endpoints) + 1}"] = x
            prev = x
        return endpoints


class Up(nn.Module):
    """bilinear upsample -> cat([skip, up]) -> 2 x (conv3x3 no-bias + BN + ReLU)."""

    def __init__(self, inC, outC, scale_factor=2, norm_layer=nn.BatchNorm2d):
        super().__init__()
        self.up = nn.Upsample(scale_factor=scale_factor, mode="bilinear", align_corners=False)
        self.conv = nn.Sequential(
            nn.Conv2d(inC, outC, kernel_size=3, padding=1, bias=False), norm_layer(outC),
            nn.ReLU(inplace=True),
            nn.Conv2d(outC, outC, kernel_size=3, padding=1, bias=False), norm_layer(outC),
            nn.ReLU(inplace=True))
        object.__setattr__(self, "_f0", FusedConv(self.conv[0], self.conv[1]))
        object.__setattr__(self, "_f1", FusedConv(self.conv[3], self.conv[4]))

    def forward_train(self, x1, x2):
        x = ag.UpCatFn.apply(x1, x2, self.up.scale_factor)
        x = ag.bn_act(ag.conv2d(x, self.conv[0]), self.conv[1], "relu")
        return ag.bn_act(ag.conv2d(x, self.conv[3]), self.conv[4], "relu")

    def upcat_nhwc(self, x1, x2):
        """cat([x2, bilinear(x1)]) as the first conv's input (its 3xFP16 operand when that conv runs on the tensor
        cores).  Separate from forward_nhwc so that sibling blocks fed the SAME (x1, x2) -- the three DeconvHeads of the
        BEV decoder -- share one up-sampling pass."""
        sf = self.up.scale_factor
        sh, sw = (sf, sf) if not isinstance(sf, (tuple, list)) else sf
        Ho, Wo = int(math.floor(x1.shape[1] * sh)), int(math.floor(x1.shape[2] * sw))
        return engine.upsample_concat_for(self._f0, x2, x1, (Ho, Wo), sf)

    def forward_nhwc(self, x1, x2, split_out=None, upcat=None):
        """split_out ("only" | "both" | None): how the consumer wants the block's output (engine.FusedConv.__call__).
        upcat: the result of a sibling's upcat_nhwc(x1, x2) with the same channel counts and scale factor."""
        if self.training:
            return self.forward_train(x1, x2)
        x = self.upcat_nhwc(x1, x2) if upcat is None else upcat
        Ho, Wo = x.shape[1], x.shape[2]
        # conv -> BN -> ReLU -> conv: the intermediate has one consumer, so the first conv's epilogue writes it
        # directly as the second conv's 3xFP16 operand (no fp32 tensor, no split pre-pass)
        mid = (x.shape[0], Ho, Wo, self.conv[0].weight.shape[0])
        return self._f1(self._f0(x, act="relu", split_out="only" if self._f1.split_ok(mid) else None), act="relu",
                        split_out=split_out)

    def forward(self, x1, x2):
        y = self.forward_nhwc(ops.nchw_to_nhwc(x1.float()), ops.nchw_to_nhwc(x2.float()))
        return ops.nhwc_to_nchw(y)


class EffNet(nn.Module):
    def __init__(self, name, inC, outC, image_size, downsample, return_2nd_last_layer_output=True,
                 apply_final_batch_norm=False):
        super().__init__()
        if name != "efficientnet-b0":
            raise NotImplementedError
        self.trunk = EfficientNetB0(in_channels=inC)
        channels = [320, 112, 40, 24, 16, inC]
        scaled = None
        if image_size is not None:
            scaled = [tuple(image_size)]
            for _ in range(5):
                scaled.insert(0, (scaled[0][0] // 2, scaled[0][1] // 2))
        scale, i, C = 32 // downsample, 0, channels[0]
        while scale > 1:
            if scaled is None or not (scaled[i + 1][0] % 2 or scaled[i + 1][1] % 2):
                sf = 2
            else:
                sf = (scaled[i + 1][0] / scaled[i][0], scaled[i + 1][1] / scaled[i][1])
            scale //= 2
            i += 1
            C += channels[i]
            setattr(self, f"up{i}", Up(C, C, sf))
        self.n_ups = i
        self.conv = nn.Conv2d(C, outC, kernel_size=1, padding=0)
        if apply_final_batch_norm:
            self.bn = nn.BatchNorm2d(outC)
        self.apply_final_batch_norm = apply_final_batch_norm
        self.return_2nd_last_layer_output = return_2nd_last_layer_output
        object.__setattr__(self, "_f_out", FusedConv(self.conv, self.bn if apply_final_batch_norm
                                                     else None))

    def forward_train(self, x):
        ep = self.trunk.extract_endpoints_train(x)
        y = ep["reduction_5"]
        for i in range(1, self.n_ups + 1):
            y = getattr(self, f"up{i}").forward_train(y, ep[f"reduction_{5 - i}"])
        out = ag.conv2d(y, self.conv)
        if self.apply_final_batch_norm:
            out = ag.bn_act(out, self.bn, "relu")
        return (out, y) if self.return_2nd_last_layer_output else out

    def forward_nhwc(self, x):
        if self.training:
            return self.forward_train(x)
        ep = self.trunk.extract_endpoints_nhwc(x)
        n = 5
        y = ep[f"reduction_{n}"]
        for i in range(1, self.n_ups + 1):
            skip = ep[f"reduction_{n - i}"]
            so = None
            if i == self.n_ups and not self.return_2nd_last_layer_output:
                # the last block's output only feeds the final 1x1 conv
                up = getattr(self, f"up{i}")
                sf = up.up.scale_factor
                sh, sw = (sf, sf) if not isinstance(sf, (tuple, list)) else sf
                shp = (y.shape[0], int(math.floor(y.shape[1] * sh)), int(math.floor(y.shape[2] * sw)),
                       up.conv[3].weight.shape[0])
                so = "only" if self._f_out.split_ok(shp) else None
            y = getattr(self, f"up{i}").forward_nhwc(y, skip, split_out=so)
        # the features feed tensor-core convs (depth head, distillation head) AND fp32 readers: write both forms
        out = self._f_out(y, act="relu" if self.apply_final_batch_norm else "none", split_out="both")
        return (out, y) if self.return_2nd_last_layer_output else out

    def forward(self, x):
        r = self.forward_nhwc(ops.nchw_to_nhwc(x.float()))
        if isinstance(r, tuple):
            return tuple(ops.nhwc_to_nchw(t) for t in r)
        return ops.nhwc_to_nchw(r)


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/blocks/effnet.py")
