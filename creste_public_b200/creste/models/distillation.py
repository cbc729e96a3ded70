"""Mirror of reference creste/models/distillation.py (DistillationBackbone :18-207)."""
import os

import torch
from torch import nn

from creste_public_b200 import autograd as ag
from creste_public_b200 import ops
from creste_public_b200.engine import require_eval
from .blocks.conv import MultiLayerConv  # noqa: F401  (resolved by name through globals())
from .depth import DepthCompletion  # noqa: F401


class DistillationBackbone(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.vision_cfg = model_cfg.vision_backbone
        self.depth_cfg = model_cfg.depth_head
        self.distillation_cfg = model_cfg.distillation_head
        self.input_image_shape = self.vision_cfg["effnet_cfgs"]["image_size"]
        self.ckpt_path = self.model_cfg.get("ckpt_path", "") or self.vision_cfg.get("ckpt_path", "")
        self.multiview_distillation = self.model_cfg.get("multiview_distillation", False)
        self.weights_path = self.model_cfg.get("weights_path", "") or \
            self.vision_cfg.get("weights_path", "")
        self.freeze_weights = model_cfg.get("freeze_weights", False)
        self.depth_trunk = self.vision_cfg.get("depth_trunk", "DepthCompletion")
        if self.depth_trunk not in globals():
            raise NotImplementedError(f"Model {self.depth_trunk} not found")
        self.depthcomp = globals()[self.depth_trunk](self.model_cfg)
        if self.multiview_distillation:
            raise NotImplementedError("multiview_distillation is disabled in every shipped config "
                                      "(configs/model/*/ *.yaml) and is not on the hot path")
        self.pe_map_cfg = model_cfg.get("pe_map", None)
        if self.pe_map_cfg is not None:
            raise NotImplementedError("pe_map is unused in the shipped configs")
        head_cfg = self.distillation_cfg.feature_head
        if head_cfg.name not in globals():
            raise NotImplementedError(f"Feature head {head_cfg.name} not found")
        self.dino_head = globals()[head_cfg.name](head_cfg)
        if os.path.isfile(self.ckpt_path):
            self.load_weights(self.ckpt_path)
        if os.path.isfile(self.weights_path):
            self.load_weights(self.weights_path)

    def load_weights(self, weights_path):
        """Key surgery of reference distillation.py:94-127."""
        sd = torch.load(weights_path, weights_only=False)["state_dict"]
        sd = {k.replace("model.", "", 1): v for k, v in sd.items() if k.startswith("model.")}
        sd = {k.replace("depthcomp.depthcomp.", "depthcomp.", 1): v for k, v in sd.items()}
        sd = {k.replace("depthcomp.dino_head.", "dino_head.", 1): v for k, v in sd.items()}
        sd = {k: v for k, v in sd.items() if "bevclassifier" not in k and "cam2map" not in k}
        self.load_state_dict(sd, strict=True)
        if self.freeze_weights:
            for name, p in self.named_parameters():
                p.requires_grad = name not in sd

    def unfreeze_backbone(self):
        for p in self.depthcomp.parameters():
            p.requires_grad = True

    def forward_nhwc(self, x_nhwc, B, V, want_nchw=True, want_dino=True):
        out, nh = self.depthcomp.forward_nhwc(x_nhwc, want_nchw)
        if want_dino:
            d = self.dino_head.forward_nhwc(nh["feats"])
            BV, Hs, Ws, D = d.shape
            to_nchw = ag.ToNCHW.apply if d.requires_grad else ops.nhwc_to_nchw
            out["dino_pe_feats"] = to_nchw(d).view(B, 1, D, Hs, Ws)
        return out, nh

    def forward(self, x):
        rgbd = x
        B, V, Cc, H, W = rgbd.shape
        out, _ = self.forward_nhwc(ops.nchw_to_nhwc(rgbd.reshape(B * V, Cc, H, W).float()), B, V)
        return out


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/distillation.py")
