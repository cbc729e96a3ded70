"""Mirror of reference creste/models/terrainnet.py (TerrainNet :24-350): RGB-D backbone ->
frustum->BEV splat -> BEV decoder.  Weight-loading modes and key surgery follow :111-261."""
import os

import torch
from torch import nn

from creste_public_b200 import ops
from creste_public_b200.config import OmegaConf
from creste_public_b200.engine import require_eval
from .blocks.inpainting import InpaintingResNet18MultiHead  # noqa: F401  (globals() lookup)
from .blocks.splat_projection import Camera2MapMulti
from .depth import DepthCompletion  # noqa: F401
from .distillation import DistillationBackbone  # noqa: F401


class TerrainNet(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.views = model_cfg.get("views", 1)
        self.vision_cfg = model_cfg.vision_backbone
        self.camproj_cfg = model_cfg.camera_projector
        self.depth_cfg = model_cfg.depth_head
        self.discretize_cfg = model_cfg.discretize
        self.projector_cfg = model_cfg.get("projection_head", None)
        self.ckpt_path = model_cfg.get("ckpt_path", "")
        self.weights_path = model_cfg.get("weights_path", "")
        self.freeze_weights = model_cfg.get("freeze_weights", False)
        self.use_temporal = model_cfg.get("use_temporal", False)
        self.use_movability = model_cfg.get("use_movability", False)
        self.load_setting = model_cfg.get("load_setting", "strict")
        self.drop_decoder = model_cfg.get("drop_decoder", False)
        self.bev_classifer_cfg = model_cfg.get("bev_classifier", None)
        if self.bev_classifer_cfg is not None:
            self.bev_classifer_cfg = OmegaConf.to_object(model_cfg.bev_classifier)
        self.bev_semantic_head_cfg = model_cfg.get("bev_semantic_head", None)
        if self.bev_semantic_head_cfg is not None:
            raise NotImplementedError("bev_semantic_head is not used by the shipped configs")
        if self.use_temporal:
            raise NotImplementedError("use_temporal (ConvGRU) is False in every shipped config")
        name = self.vision_cfg.get("class_name", None) or "DistillationBackbone"
        if name not in globals():
            raise NotImplementedError(f"Vision backbone {name} not implemented")
        self.depthcomp = globals()[name](self.model_cfg)
        self.cam2map = Camera2MapMulti(self.camproj_cfg, mode="bilinear")
        self.splat_key = self.camproj_cfg.get("splat_key", "depth_preds_feats")
        self.bevclassifier = None
        if self.bev_classifer_cfg is not None:
            cname = self.bev_classifer_cfg["name"]
            if cname not in globals():
                raise NotImplementedError(f"Bev classifier {cname} not implemented")
            self.bevclassifier = globals()[cname](**self.bev_classifer_cfg["net_kwargs"])
        if os.path.isfile(self.weights_path) and not os.path.isfile(self.ckpt_path):
            self.load_weights(self.weights_path)

    # ------------------------------------------------------------------ weights (terrainnet.py:111-261)
    def load_weights(self, weights_path):
        sd = torch.load(weights_path, weights_only=False)["state_dict"]
        sd = {(k.replace("model.", "", 1) if k.startswith("model.") else k): v for k, v in sd.items()}
        n0 = len(sd)
        # stage-1 checkpoints store the backbone as depthcomp.* / dino_head.*; here it lives one
        # level down (self.depthcomp is the DistillationBackbone)
        fixed = {}
        for k, v in sd.items():
            if k.startswith("depthcomp.") and not k.startswith("depthcomp.depthcomp.") and \
                    not k.startswith("depthcomp.dino_head."):
                k = "depthcomp." + k
            elif k.startswith("dino_head."):
                k = "depthcomp." + k
            fixed[k] = v
        assert len(fixed) == n0
        sd = fixed
        setting = self.load_setting

        def trainable_iff(pred):
            for name, p in self.named_parameters():
                p.requires_grad = bool(pred(name))

        heads = "bevclassifier.out_heads"
        if setting in ("strict", "strict_freeze", "strict_unfreezesplat"):
            # Lightning stage-2 checkpoints carry the LossManager's tensors under `loss.`: dropped in all
            # three strict modes (terrainnet.py:236-257); only strict_unfreezesplat loads non-strictly
            sd = {k: v for k, v in sd.items() if not k.startswith("loss.")}
            self.load_state_dict(sd, strict=(setting != "strict_unfreezesplat"))
            if setting == "strict_freeze":
                trainable_iff(lambda n: False)
            elif setting == "strict_unfreezesplat":
                trainable_iff(lambda n: "cam2map." in n)
        elif setting == "ft_semantic_head":
            # everything loads non-strictly; only a (future) semantic head and the 1-channel decoder
            # heads train (terrainnet.py:151-166)
            self.load_state_dict(sd, strict=False)
            trainable_iff(lambda n: "bev_semantic_head" in n)
            for head in self.bevclassifier.out_heads:
                if head.proj.out_channels == 1:
                    for p in head.parameters():
                        p.requires_grad = True
        elif setting == "ft_decoders_all":
            sd = {k: v for k, v in sd.items() if heads not in k}
            self.load_state_dict(sd, strict=False)
            trainable_iff(lambda n: heads in n)
        elif setting == "ft_decoders_partial":
            last = lambda n: heads in n and ("up2" in n or "proj" in n)   # noqa: E731
            sd = {k: v for k, v in sd.items() if not last(k)}
            self.load_state_dict(sd, strict=False)
            trainable_iff(last)
        else:
            raise ValueError(f"Invalid load_setting {self.load_setting}")

    # ------------------------------------------------------------------ forward (terrainnet.py:272-350)
    def forward_full(self, x, want_nchw=True, want_dino=True):
        """Returns (reference-layout dict, preds NHWC per head prefix).  Follows self.training like the
        reference: eval = the fused inference engine (running statistics); train = the autograd graph over the
        same kernels (BatchNorm batch statistics through backbone, splat and BEV decoder; stage 2)."""
        if self.training and self.use_movability:
            raise NotImplementedError("use_movability (double splat with movability masks) is False in every "
                                      "shipped config")
        rgbd, p2p = x[:2]
        B, N, Cc, H, W = rgbd.shape
        x_nhwc = ops.nchw_to_nhwc(rgbd.reshape(B * N, Cc, H, W).float())
        outputs, nh = self.depthcomp.forward_nhwc(x_nhwc, B, N, want_nchw, want_dino)
        if self.splat_key != "depth_preds_feats":
            raise NotImplementedError("splat_key other than depth_preds_feats")
        depth = outputs["depth_preds_metric"]                     # [B*N, Hs, Ws]
        ret, bev_nhwc = self.cam2map.forward_nhwc(depth, nh["feats"], p2p.reshape(B * N, 4, 4).float(),
                                                  want_nchw)
        outputs.update(ret)
        preds_nhwc = None
        if self.bevclassifier is not None:
            if self.bevclassifier.input_key != "bev_features":
                raise NotImplementedError("bev_classifier.input_key must be bev_features")
            dec, preds_nhwc = self.bevclassifier.forward_nhwc(bev_nhwc, want_nchw)
            outputs.update(dec)
        return outputs, preds_nhwc

    def forward(self, x):
        return self.forward_full(x)[0]


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/terrainnet.py")
