"""Mirror of reference creste/models/depth.py (DepthCompletion :17-158)."""
import os

import torch
from torch import nn

from creste_public_b200 import autograd as ag
from creste_public_b200 import ops
from creste_public_b200.engine import require_eval
from .blocks.conv import MultiLayerConv
from .vision_encoder import VisionEncoder


class DepthCompletion(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.vision_cfg = model_cfg.vision_backbone
        self.depth_cfg = model_cfg.depth_head
        self.discretize_cfg = model_cfg.discretize
        self.return_feats = self.vision_cfg.return_feats
        self.vision_backbone = VisionEncoder(self.vision_cfg)
        self.depth_head = MultiLayerConv(self.depth_cfg)
        if os.path.isfile(self.vision_cfg.weights_path):
            self.load_weights(self.vision_cfg.weights_path)

    def load_weights(self, weights_path):
        """Lightning checkpoint -> this module (key surgery of reference depth.py:34-58)."""
        sd = torch.load(weights_path, weights_only=False)["state_dict"]
        sd = {k.replace("model.", "", 1): v for k, v in sd.items() if k.startswith("model.")}
        sd = {(k.replace("depthcomp.", "", 1) if k.startswith("depthcomp.") else k): v
              for k, v in sd.items()}
        own = set(self.state_dict().keys())
        self.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=True)

    @staticmethod
    def _convert_to_metric_depth(x, discretize_cfg, valid_thres=0.9):
        """x: depth logits NCHW [B,D,H,W] -> (metric depth [B,H,W] in metres, arg-max bins)."""
        assert discretize_cfg.mode == "UD"
        m, b = ops.depth_expectation(ops.nchw_to_nhwc(x.float()), float(discretize_cfg.depth_min),
                                     float(discretize_cfg.depth_max))
        return m, b

    def forward_nhwc(self, x_nhwc, want_nchw=True):
        """Returns (outputs dict in the reference's NCHW layouts, nhwc dict for the parents)."""
        feats = self.vision_backbone.forward_nhwc(x_nhwc)
        logits = self.depth_head.forward_nhwc(feats)
        dmin, dmax = float(self.discretize_cfg.depth_min), float(self.discretize_cfg.depth_max)
        metric, bins = ops.depth_expectation(logits.detach(), dmin, dmax)
        if logits.requires_grad and torch.is_grad_enabled():
            # stage 2 differentiates the soft-argmax depth (SmoothL1Depth on depth_preds_metric and the splat's
            # voxel coordinates, train_ssc.py:92-129); the arg-max bins stay a value
            metric = ag.DepthExpectFn.apply(logits, dmin, dmax, 1000.0)
        out = {"depth_preds_metric": metric, "depth_preds_bins": bins}
        if want_nchw:
            to_nchw = ag.ToNCHW.apply if logits.requires_grad else ops.nhwc_to_nchw
            out["depth_preds_logits"] = to_nchw(logits)
            if self.return_feats:
                out["depth_preds_feats"] = to_nchw(feats)
        return out, {"feats": feats, "logits": logits}

    def forward(self, x):
        out, _ = self.forward_nhwc(ops.nchw_to_nhwc(x.float()))
        return out


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/depth.py")
