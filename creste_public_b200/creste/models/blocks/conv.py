"""Config-driven conv stacks, mirror of reference creste/models/blocks/conv.py:
MultiLayerConv (:5-32), ConvEncoder (:37-58), ConvLayer (:63-85), MultiScaleFCN (:88-161).

The nn.Conv2d / nn.BatchNorm2d members are parameter containers with the reference's names
(so checkpoints load unchanged); compute goes through creste_conv2d (include/creste_b200.h)
with BatchNorm folded into the conv epilogue.  Public `forward` takes / returns NCHW like the
reference; `forward_nhwc` is the layout-native entry used by the parent modules.
"""
import torch
from torch import nn

from creste_public_b200 import ops
from creste_public_b200.engine import FusedConv, require_eval


class _ConvStack(nn.Module):
    """[conv (+BN) + ReLU] x n held in one nn.Sequential under `attr`."""

    def _build(self, dims, kernels, paddings, strides, norm):
        m = []
        for i, k in enumerate(kernels):
            m.append(nn.Conv2d(dims[i], dims[i + 1], k, padding=paddings[i], stride=strides[i]))
            if norm == "batch_norm":
                m.append(nn.BatchNorm2d(dims[i + 1]))
            m.append(nn.ReLU())
        return nn.Sequential(*m)

    def _fused(self, seq):
        if not hasattr(self, "_fc"):
            layers = list(seq)
            fc = []
            i = 0
            while i < len(layers):
                conv = layers[i]
                bn = layers[i + 1] if isinstance(layers[i + 1], nn.BatchNorm2d) else None
                # layers after the first are fed a pre-split operand by their producer's epilogue (forward_nhwc):
                # with no pre-pass to pay, 1x1 convs with 64 <= C <= 192 also belong on the tensor cores
                fc.append(FusedConv(conv, bn, prefer_tc=len(fc) > 0))
                i += 3 if bn is not None else 2
            object.__setattr__(self, "_fc", fc)
        return self._fc

    def forward_train(self, x):
        """Train-mode stack as an autograd graph: conv (+bias) -> BatchNorm with batch statistics ->
        ReLU per layer (reference conv.py:5-32 under nn.Module.train())."""
        from creste_public_b200 import autograd as ag
        layers = list(self._seq())
        i = 0
        while i < len(layers):
            conv = layers[i]
            bn = layers[i + 1] if isinstance(layers[i + 1], nn.BatchNorm2d) else None
            y = ag.conv2d(x, conv)
            x = ag.bn_act(y, bn, "relu") if bn is not None else ag.relu(y)
            i += 3 if bn is not None else 2
        return x

    def forward_nhwc(self, x):
        if self.training:
            return self.forward_train(x)
        fs = self._fused(self._seq())
        for i, f in enumerate(fs):
            so = None
            if i + 1 < len(fs):
                # conv (+BN) + ReLU -> conv: the intermediate has one consumer; write it as that conv's operand
                N, H, W = x.shape[0], x.shape[1], x.shape[2]
                kh, kw = f.conv.kernel_size
                ph, pw = f.conv.padding if isinstance(f.conv.padding, tuple) else (f.conv.padding,) * 2
                st = f.conv.stride[0] if isinstance(f.conv.stride, tuple) else f.conv.stride
                mid = (N, (H + 2 * ph - kh) // st + 1, (W + 2 * pw - kw) // st + 1, f.conv.out_channels)
                so = "only" if fs[i + 1].split_ok(mid) else None
            x = f(x, act="relu", split_out=so)
        return x

    def forward(self, x):
        return ops.nhwc_to_nchw(self.forward_nhwc(ops.nchw_to_nhwc(x.float())))


class MultiLayerConv(_ConvStack):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.kernels, self.paddings, self.dims = model_cfg.kernels, model_cfg.paddings, model_cfg.dims
        self.norm_type = model_cfg.norm_type
        self.stride = model_cfg.get("stride", [1] * len(self.kernels))
        self.model = self._build(self.dims, self.kernels, self.paddings, self.stride, self.norm_type)

    def _seq(self):
        return self.model


class ConvEncoder(_ConvStack):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        k, p = model_cfg["kernels"], model_cfg["paddings"]
        assert len(k) == len(p)
        self.convs = self._build(model_cfg["dims"], k, p, [1] * len(k), model_cfg["norm_type"])

    def _seq(self):
        return self.convs


class ConvLayer(nn.Sequential):
    """conv(k, pad=k//2, bias) [+BN] [+ReLU] with members named conv / norm / relu."""

    def __init__(self, in_channels, out_channels, kernel=3, stride=1, dropout=0.1, bn=False,
                 norm_type="batch_norm", relu=True, bias=False):
        super().__init__()
        self.add_module("conv", nn.Conv2d(in_channels, out_channels, kernel_size=kernel,
                                          stride=stride, padding=kernel // 2, bias=bias))
        if bn:
            if norm_type != "batch_norm":
                raise Exception("Unknown norm type:", norm_type)
            self.add_module("norm", nn.BatchNorm2d(out_channels))
        if relu:
            self.add_module("relu", nn.ReLU(inplace=True))

    def fused(self):
        if "_fused" not in self.__dict__:
            object.__setattr__(self, "_fused", FusedConv(self.conv, getattr(self, "norm", None)))
        return self.__dict__["_fused"]

    def forward_nhwc(self, x):
        return self.fused()(x, act="relu" if hasattr(self, "relu") else "none")

    def forward_autograd_nhwc(self, x):
        """Differentiable path (training / loss evaluation): conv -> BN (batch statistics when
        self.norm.training, as F.batch_norm) -> ReLU, composed of creste_public_b200.autograd
        Functions so that first- and second-order gradients flow."""
        from creste_public_b200 import autograd as ag
        y = ag.conv2d(x, self.conv)
        has_relu = hasattr(self, "relu")
        if hasattr(self, "norm"):
            return ag.batch_norm(y, self.norm, relu=has_relu)
        return ag.relu(y) if has_relu else y

    def forward(self, x):
        if _wants_grad(self, x):
            from creste_public_b200 import autograd as ag
            return ag.ToNCHW.apply(self.forward_autograd_nhwc(ag.ToNHWC.apply(x.float())))
        if self.training and hasattr(self, "norm"):       # batch statistics without a graph
            return ops.nhwc_to_nchw(self.forward_autograd_nhwc(ops.nchw_to_nhwc(x.float())))
        return ops.nhwc_to_nchw(self.forward_nhwc(ops.nchw_to_nhwc(x.float())))


def _wants_grad(module, x):
    """True when the reference's autograd would record this call: grad mode on and either the
    input or a parameter of the module requires grad."""
    if not torch.is_grad_enabled():
        return False
    return bool(x.requires_grad) or any(p.requires_grad for p in module.parameters())


class MultiScaleFCN(nn.Module):
    """Reward net (reference conv.py:88-161): prepool -> {skip || maxpool -> trunk -> bilinear x2}
    -> cat -> postpool.  The trunk's `ConvLayer(relu) -> BN -> ReLU` triple is evaluated as the
    reference does: ReLU(conv) first, then BN as a per-channel affine, then ReLU."""

    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.prepool_cfg, self.postpool_cfg = model_cfg.prepool, model_cfg.postpool
        self.skip_cfg, self.trunk_cfg = model_cfg.skip, model_cfg.trunk

        def stack(cfg):
            return nn.Sequential(*[
                ConvLayer(cfg.dims[i], cfg.dims[i + 1], kernel=cfg.kernels[i], stride=cfg.stride[i],
                          bn=True, norm_type=cfg.norm_type, relu=True, bias=False)
                for i in range(len(cfg.kernels))])

        self.prepool = stack(self.prepool_cfg)
        self.skip = stack(self.skip_cfg)
        trunk = [nn.MaxPool2d(kernel_size=2, stride=2)]
        for i in range(len(self.trunk_cfg.kernels)):
            trunk.append(ConvLayer(self.trunk_cfg.dims[i], self.trunk_cfg.dims[i + 1],
                                   kernel=self.trunk_cfg.kernels[i]))
            if self.trunk_cfg.norm_type == "batch_norm":
                trunk.append(nn.BatchNorm2d(self.trunk_cfg.dims[i + 1]))
            trunk.append(nn.ReLU(inplace=True))
        trunk.append(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
        self.trunk = nn.Sequential(*trunk)
        self.postpool = stack(self.postpool_cfg)
        self.initialize_weights_with_xavier()

    def initialize_weights_with_xavier(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def _trunk_affines(self):
        """BN layers that follow a ReLU'd ConvLayer in the trunk, as 1x1 diagonal 'convs'."""
        if "_taff" not in self.__dict__:
            object.__setattr__(self, "_taff", {})
        return self.__dict__["_taff"]

    def forward_nhwc(self, x):
        if self.training:
            return self.forward_autograd_nhwc(x)
        from creste_public_b200.engine import PackCache, bn_scale_shift
        for l in self.prepool:
            x = l.forward_nhwc(x)
        skip = x
        for l in self.skip:
            skip = l.forward_nhwc(skip)
        t = ops.maxpool2_concat([x])
        layers = list(self.trunk)[1:-1]
        i = 0
        while i < len(layers):
            cl = layers[i]
            t = cl.forward_nhwc(t)                     # ReLU(conv(t))
            if i + 1 < len(layers) and isinstance(layers[i + 1], nn.BatchNorm2d):
                bn = layers[i + 1]
                cache = self._trunk_affines().setdefault(i, PackCache())
                Cc = bn.num_features

                def build(bn=bn, Cc=Cc):
                    scale, shift = bn_scale_shift(bn)
                    eye = torch.eye(Cc, device=scale.device).view(Cc, Cc, 1, 1)
                    return ops.pack_conv_weight(eye), scale, shift
                w, scale, shift = cache.get("aff", [bn.weight, bn.bias, bn.running_mean,
                                                     bn.running_var], build)
                # BN + ReLU as an exact per-channel affine (identity 1x1 conv: products by 1/0
                # are exact, so this equals the reference's elementwise BN)
                t = ops.conv2d(t, w, Cc, 1, 1, 1, (0, 0, 0, 0), scale, shift, None, None, "relu",
                               False, "fp32")
                i += 3
            else:
                i += 2
        N, H, W, _ = skip.shape
        cat = ops.upsample_concat(skip, t, (H, W), 2, x_first=True)   # cat([up(trunk), skip])
        for l in self.postpool:
            cat = l.forward_nhwc(cat)
        return cat

    def forward_autograd_nhwc(self, x):
        """Differentiable (to second order) evaluation of the reward FCN; BatchNorm layers honour
        their own .training flag like the reference (batch statistics + running-stat update in
        train mode).  Reference graph: conv.py:148-161."""
        from creste_public_b200 import autograd as ag
        for l in self.prepool:
            x = l.forward_autograd_nhwc(x)
        skip = x
        for l in self.skip:
            skip = l.forward_autograd_nhwc(skip)
        t = x
        for m in self.trunk:
            if isinstance(m, nn.MaxPool2d):
                t = ag.MaxPool2Fn.apply(t)
            elif isinstance(m, ConvLayer):
                t = m.forward_autograd_nhwc(t)
            elif isinstance(m, nn.BatchNorm2d):
                t = ag.batch_norm(t, m, relu=False)
            elif isinstance(m, nn.ReLU):
                t = ag.relu(t)
            elif isinstance(m, nn.Upsample):
                t = ag.Up2Fn.apply(t)
            else:
                raise NotImplementedError(type(m).__name__)
        cat = torch.cat([t, skip], dim=-1)
        for l in self.postpool:
            cat = l.forward_autograd_nhwc(cat)
        return cat

    def forward(self, x):
        """Expects input of shape [B, C, H, W] (NCHW), returns [B, 1, H, W]."""
        if _wants_grad(self, x):
            from creste_public_b200 import autograd as ag
            return ag.ToNCHW.apply(self.forward_autograd_nhwc(ag.ToNHWC.apply(x.float())))
        return ops.nhwc_to_nchw(self.forward_nhwc(ops.nchw_to_nhwc(x.float())))


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/blocks/conv.py")
