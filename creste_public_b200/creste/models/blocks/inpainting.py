"""BEV decoder: mirror of reference creste/models/blocks/inpainting.py (Inpainting :8-50,
DeconvHead :52-68, InpaintingResNet18MultiHead :70-109).  The ResNet-18 layers are plain
parameter containers with torchvision's names and initialisation; all compute is
creste_conv2d (BN folded, residual + ReLU fused) and creste_upsample_concat.
"""
import torch
from torch import nn

from creste_public_b200 import ops
from creste_public_b200.engine import FusedConv, carry_amax, require_eval, upsample_concat_for
from .effnet import Up


def prefix_dict(prefix, d, seprator="/"):
    return {prefix + seprator + k: v for k, v in d.items()}


class BasicBlock(nn.Module):
    """torchvision.models.resnet.BasicBlock parameter layout (conv1/bn1/conv2/bn2/downsample)."""

    def __init__(self, inplanes, planes, stride, norm_layer):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = norm_layer(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False),
                                            norm_layer(planes))
        self.stride = stride
        object.__setattr__(self, "_f1", FusedConv(self.conv1, self.bn1))
        object.__setattr__(self, "_f2", FusedConv(self.conv2, self.bn2))
        if self.downsample is not None:
            object.__setattr__(self, "_fd", FusedConv(self.downsample[0], self.downsample[1]))

    def forward_train(self, x):
        """torchvision BasicBlock.forward in train mode (BatchNorm batch statistics) as an autograd graph."""
        from creste_public_b200 import autograd as ag
        out = ag.bn_act(ag.conv2d(x, self.conv1), self.bn1, "relu")
        out = ag.bn_act(ag.conv2d(out, self.conv2), self.bn2, "none")
        idt = x if self.downsample is None else ag.bn_act(ag.conv2d(x, self.downsample[0]), self.downsample[1], "none")
        return ag.relu(ag.AddScaledFn.apply(out, idt, None))

    def forward_nhwc(self, x):
        if self.training:
            return self.forward_train(x)
        idt = self._fd(x) if self.downsample is not None else x
        # conv1's output has one consumer (conv2): its epilogue writes conv2's 3xFP16 operand directly
        N, H, W, _ = x.shape
        mid = (N, (H - 1) // self.stride + 1, (W - 1) // self.stride + 1, self.conv1.weight.shape[0])
        return self._f2(self._f1(x, act="relu", split_out="only" if self._f2.split_ok(mid) else None), act="relu",
                        residual=idt)


def _make_layer(inplanes, planes, stride, norm_layer):
    return nn.Sequential(BasicBlock(inplanes, planes, stride, norm_layer),
                         BasicBlock(planes, planes, 1, norm_layer))


class Inpainting(nn.Module):
    def __init__(self, input_key=None, output_prefix=None, learnable_loss_weight=False):
        super().__init__()
        self.input_key = input_key or "merged_bev_features"
        self.output_prefix = output_prefix or "inpainting"
        self.log_var = nn.Parameter(torch.tensor([0.0]), requires_grad=True) \
            if learnable_loss_weight else None

    def _collect(self, out, key_suffix=""):
        if isinstance(out, list):
            assert isinstance(self.output_prefix, list) and len(out) == len(self.output_prefix)
            ret = {}
            for p, o in zip(self.output_prefix, out):
                if p == "inpainting_sam":
                    p = f"{p}{key_suffix}"
                ret.update(prefix_dict(p, o, seprator="_"))
            return ret
        return prefix_dict(f"{self.output_prefix}{key_suffix}", out, seprator="_")

    def forward(self, tensor_dict, key_suffix=""):
        x = tensor_dict[f"{self.input_key}{key_suffix}"]
        out = self._forward(x)
        if self.log_var is not None:
            out["log_variance"] = self.log_var
        return self._collect(out, key_suffix)


class DeconvHead(nn.Module):
    def __init__(self, in_ch, out_ch, norm_layer):
        super().__init__()
        self.up1 = Up(in_ch, 256, scale_factor=4, norm_layer=norm_layer)
        self.up2 = nn.Sequential(
            nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
            nn.Conv2d(256, 128, kernel_size=3, padding=1, bias=False), norm_layer(128),
            nn.ReLU(inplace=True))
        self.proj = nn.Conv2d(128, out_ch, kernel_size=1, padding=0)
        object.__setattr__(self, "_f_up2", FusedConv(self.up2[1], self.up2[2]))
        object.__setattr__(self, "_f_proj", FusedConv(self.proj, None))

    def forward_train(self, x1, x2):
        from creste_public_b200 import autograd as ag
        x = self.up1.forward_train(x1, x2)
        x = ag.bn_act(ag.conv2d(ag.Up2Fn.apply(x), self.up2[1]), self.up2[2], "relu")
        return ag.conv2d(x, self.proj), x

    def forward_nhwc(self, x1, x2):
        if self.training:
            return self.forward_train(x1, x2)
        x = self.up1.forward_nhwc(x1, x2)
        N, H, W, _ = x.shape
        x = upsample_concat_for(self._f_up2, None, x, (2 * H, 2 * W), 2)
        x = self._f_up2(x, act="relu")
        return self._f_proj(x, act="none"), x

    def forward_fused_nhwc(self, x1, x2, want_nchw=True, upcat=None):
        """Eval path with the projection head, the NCHW copy of the predictions and the NCHW copy of the 128-channel
        features in ONE pass over the features (creste_proj_head) instead of a generic conv + two transposes.
        -> (pred NHWC, features NHWC, pred NCHW | None, features NCHW | None).  upcat: see Up.forward_nhwc."""
        x = self.up1.forward_nhwc(x1, x2, upcat=upcat)
        N, H, W, _ = x.shape
        x = upsample_concat_for(self._f_up2, None, x, (2 * H, 2 * W), 2)
        x = self._f_up2(x, act="relu")
        K, Cc = self.proj.weight.shape[0], self.proj.weight.shape[1]
        if K > 32 or Cc % 32 or Cc > 512:
            pred = self._f_proj(x, act="none")
            return pred, x, (ops.nhwc_to_nchw(pred) if want_nchw else None), (ops.nhwc_to_nchw(x) if want_nchw else None)
        w, b = self._f_proj.cache.get("proj_kc", [self.proj.weight, self.proj.bias], lambda: (
            self.proj.weight.detach().float().reshape(K, Cc).contiguous(),
            None if self.proj.bias is None else self.proj.bias.detach().float().contiguous()))
        pred, pred_nchw, x_nchw = ops.proj_head(x, w, b, want_nchw, want_nchw)
        return pred, x, pred_nchw, x_nchw


class InpaintingResNet18MultiHead(Inpainting):
    def __init__(self, num_input_features, num_classes, norm_layer="batch_norm", **kwargs):
        super().__init__(**kwargs)
        if norm_layer != "batch_norm":
            raise Exception("Unsupported norm layer:", norm_layer)
        norm_layer = nn.BatchNorm2d
        self.conv1 = nn.Conv2d(num_input_features, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = norm_layer(64)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = _make_layer(64, 64, 1, norm_layer)
        self.layer2 = _make_layer(64, 128, 2, norm_layer)
        self.layer3 = _make_layer(128, 256, 2, norm_layer)
        # torchvision resnet18(zero_init_residual=True) initialisation for the borrowed layers
        for m in [self.bn1, self.layer1, self.layer2, self.layer3]:
            for mm in m.modules():
                if isinstance(mm, nn.Conv2d):
                    nn.init.kaiming_normal_(mm.weight, mode="fan_out", nonlinearity="relu")
                elif isinstance(mm, nn.BatchNorm2d):
                    nn.init.constant_(mm.weight, 1)
                    nn.init.constant_(mm.bias, 0)
        for layer in (self.layer1, self.layer2, self.layer3):
            for blk in layer:
                nn.init.constant_(blk.bn2.weight, 0)
        self.out_heads = nn.ModuleList([DeconvHead(64 + 256, n, norm_layer) for n in num_classes])
        object.__setattr__(self, "_f_conv1", FusedConv(self.conv1, self.bn1))

    def forward_nhwc(self, bev_nhwc, want_nchw=True):
        """Returns (reference-layout dict, {prefix: preds NHWC}).  In train mode every node is an autograd
        Function over the sm_100a kernels (BatchNorm batch statistics, strided-conv gradients)."""
        from creste_public_b200 import autograd as ag
        if self.training:
            x = ag.bn_act(ag.conv2d(bev_nhwc, self.conv1), self.bn1, "relu")
        else:
            x = self._f_conv1(bev_nhwc, act="relu")
        x1 = x
        for blk in self.layer1:
            x1 = blk.forward_nhwc(x1)
        x = x1
        for blk in list(self.layer2) + list(self.layer3):
            x = blk.forward_nhwc(x)
        ret, preds_nhwc = {}, {}
        upcat = None
        if not self.training:
            # the heads' first stage, Up(320, 256, x4)(x, x1), sees the same two tensors in every head: one up-sampling +
            # concat pass (written once as the 3xFP16 operand) feeds the three first convs
            h0 = self.out_heads[0].up1
            same = all(h.up1.up.scale_factor == h0.up.scale_factor and
                       h.up1.conv[0].weight.shape == h0.conv[0].weight.shape for h in self.out_heads)
            if same:
                upcat = h0.upcat_nhwc(x, x1)
        for head, prefix in zip(self.out_heads, self.output_prefix):
            if not self.training:
                pred, fea, pred_nchw, fea_nchw = head.forward_fused_nhwc(x, x1, want_nchw, upcat=upcat)
                preds_nhwc[prefix] = pred
                if want_nchw:
                    ret[f"{prefix}_preds"], ret[f"{prefix}_features"] = pred_nchw, fea_nchw
                continue
            pred, fea = head.forward_nhwc(x, x1)
            preds_nhwc[prefix] = pred
            if want_nchw:
                to_nchw = ag.ToNCHW.apply if (pred.requires_grad or fea.requires_grad) else ops.nhwc_to_nchw
                ret[f"{prefix}_preds"] = to_nchw(pred)
                ret[f"{prefix}_features"] = to_nchw(fea)
        return ret, preds_nhwc

    def _forward(self, x):
        from creste_public_b200 import autograd as ag
        to_nhwc = ag.ToNHWC.apply if (x.requires_grad and torch.is_grad_enabled()) else ops.nchw_to_nhwc
        ret, _ = self.forward_nhwc(to_nhwc(x.float()))
        return [dict(preds=ret[f"{p}_preds"], features=ret[f"{p}_features"])
                for p in self.output_prefix]


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/blocks/inpainting.py")
