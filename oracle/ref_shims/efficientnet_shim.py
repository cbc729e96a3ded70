"""Restatement of lukemelas/EfficientNet-PyTorch (0.7.1), EfficientNet-B0 only.

TEST INFRASTRUCTURE (see oracle/ref_shims/__init__.py).  The reference imports this library at
creste/models/blocks/effnet.py:5 and calls `EfficientNet.from_pretrained(name)` (:37),
`set_swish(memory_efficient=False)` (:38), `utils.get_same_padding_conv2d(image_size)` (:41)
and `trunk.extract_endpoints(x)` (:83).  The library is neither vendored nor pinned by the
reference, so this file restates its published B0 algorithm from the library's public
behaviour (stage table, TF-"SAME" static padding, BN eps/momentum, SE, endpoint rule).
PARITY AT THIS BOUNDARY IS UNPINNED: the reference holds no test or golden vector for it.
`from_pretrained` would download ImageNet weights; with no network the weights are whatever
the caller loads afterwards (tests load seeded synthetic state dicts).
"""
import math
import types

import torch
from torch import nn
from torch.nn import functional as F

# (repeats, kernel, stride, expand_ratio, in_ch, out_ch), se_ratio = 0.25 everywhere
B0_STAGES = [
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
]
BN_MOMENTUM = 1 - 0.99
BN_EPS = 1e-3
DROP_CONNECT = 0.2
SE_RATIO = 0.25


def _pair(x):
    return (x, x) if isinstance(x, int) else tuple(x)


def _out_size(size, stride):
    if size is None:
        return None
    h, w = _pair(size)
    return int(math.ceil(h / stride)), int(math.ceil(w / stride))


class Conv2dStaticSamePadding(nn.Conv2d):
    """Conv with TF 'SAME' padding fixed at construction from a nominal image size."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, image_size=None, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, **kw)
        self.stride = self.stride if len(self.stride) == 2 else [self.stride[0]] * 2
        assert image_size is not None
        ih, iw = _pair(image_size)
        kh, kw_ = self.weight.shape[-2:]
        sh, sw = self.stride
        oh, ow = math.ceil(ih / sh), math.ceil(iw / sw)
        pad_h = max((oh - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0)
        pad_w = max((ow - 1) * sw + (kw_ - 1) * self.dilation[1] + 1 - iw, 0)
        if pad_h > 0 or pad_w > 0:
            self.static_padding = nn.ZeroPad2d(
                (pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2))
        else:
            self.static_padding = nn.Identity()

    def forward(self, x):
        x = self.static_padding(x)
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation,
                        self.groups)


def get_same_padding_conv2d(image_size=None):
    def make(*a, **kw):
        return Conv2dStaticSamePadding(*a, image_size=image_size, **kw)
    return make


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def drop_connect(inputs, p, training):
    if not training:
        return inputs
    keep = 1 - p
    r = keep + torch.rand([inputs.shape[0], 1, 1, 1], dtype=inputs.dtype, device=inputs.device)
    return inputs / keep * torch.floor(r)


class MBConvBlock(nn.Module):
    def __init__(self, k, s, e, cin, cout, image_size):
        super().__init__()
        self.k, self.s, self.e, self.cin, self.cout = k, s, e, cin, cout
        mid = cin * e
        if e != 1:
            self._expand_conv = get_same_padding_conv2d(image_size)(cin, mid, kernel_size=1,
                                                                    bias=False)
            self._bn0 = nn.BatchNorm2d(mid, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._depthwise_conv = get_same_padding_conv2d(image_size)(
            mid, mid, groups=mid, kernel_size=k, stride=s, bias=False)
        self._bn1 = nn.BatchNorm2d(mid, momentum=BN_MOMENTUM, eps=BN_EPS)
        nsq = max(1, int(cin * SE_RATIO))
        self._se_reduce = get_same_padding_conv2d((1, 1))(mid, nsq, kernel_size=1)
        self._se_expand = get_same_padding_conv2d((1, 1))(nsq, mid, kernel_size=1)
        self._project_conv = get_same_padding_conv2d(_out_size(image_size, s))(
            mid, cout, kernel_size=1, bias=False)
        self._bn2 = nn.BatchNorm2d(cout, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._swish = Swish()

    def set_swish(self, memory_efficient=True):
        self._swish = Swish()

    def forward(self, inputs, drop_connect_rate=None):
        x = inputs
        if self.e != 1:
            x = self._swish(self._bn0(self._expand_conv(x)))
        x = self._swish(self._bn1(self._depthwise_conv(x)))
        sq = F.adaptive_avg_pool2d(x, 1)
        sq = self._se_expand(self._swish(self._se_reduce(sq)))
        x = torch.sigmoid(sq) * x
        x = self._bn2(self._project_conv(x))
        if self.s == 1 and self.cin == self.cout:
            if drop_connect_rate:
                x = drop_connect(x, drop_connect_rate, self.training)
            x = x + inputs
        return x


class EfficientNet(nn.Module):
    def __init__(self, image_size=224, in_channels=3, num_classes=1000):
        super().__init__()
        size = _pair(image_size)
        self._conv_stem = get_same_padding_conv2d(size)(in_channels, 32, kernel_size=3, stride=2,
                                                        bias=False)
        self._bn0 = nn.BatchNorm2d(32, momentum=BN_MOMENTUM, eps=BN_EPS)
        size = _out_size(size, 2)
        blocks = []
        for (rep, k, s, e, cin, cout) in B0_STAGES:
            for r in range(rep):
                blocks.append(MBConvBlock(k, s if r == 0 else 1, e, cin if r == 0 else cout, cout,
                                          size))
                if r == 0:
                    size = _out_size(size, s)
        self._blocks = nn.ModuleList(blocks)
        self._conv_head = get_same_padding_conv2d(size)(320, 1280, kernel_size=1, bias=False)
        self._bn1 = nn.BatchNorm2d(1280, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(0.2)
        self._fc = nn.Linear(1280, num_classes)
        self._swish = Swish()

    @classmethod
    def from_pretrained(cls, model_name, **kw):
        assert model_name == "efficientnet-b0", "only B0 is restated"
        return cls()

    @classmethod
    def from_name(cls, model_name, **kw):
        return cls.from_pretrained(model_name)

    def set_swish(self, memory_efficient=True):
        self._swish = Swish()
        for b in self._blocks:
            b.set_swish(memory_efficient)

    def extract_endpoints(self, inputs):
        endpoints = {}
        x = self._swish(self._bn0(self._conv_stem(inputs)))
        prev = x
        n = len(self._blocks)
        for idx, block in enumerate(self._blocks):
            rate = DROP_CONNECT * float(idx) / n
            x = block(x, drop_connect_rate=rate)
            if prev.size(2) > x.size(2):
                endpoints[f"reduction_{len(endpoints) + 1}"] = prev
            elif idx == n - 1:
                endpoints[f"reduction_{len(endpoints) + 1}"] = x
            prev = x
        x = self._swish(self._bn1(self._conv_head(x)))
        endpoints[f"reduction_{len(endpoints) + 1}"] = x
        return endpoints


utils = types.ModuleType("efficientnet_pytorch.utils")
utils.get_same_padding_conv2d = get_same_padding_conv2d
utils.Conv2dStaticSamePadding = Conv2dStaticSamePadding
