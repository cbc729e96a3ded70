"""Mirror of reference creste/models/vision_encoder.py:8-49 (VisionEncoder)."""
from torch import nn

from creste_public_b200 import ops
from .blocks.effnet import EffNet


class VisionEncoder(nn.Module):
    def __init__(self, vision_cfg):
        super().__init__()
        self.vision_cfg = vision_cfg
        self.input_type = vision_cfg.input_type
        self.name = vision_cfg.name
        if self.input_type in ("rgb", "rgbd"):
            if "efficientnet" in self.name:
                c = self.vision_cfg.effnet_cfgs
                self.model = EffNet(name=self.name, inC=c.in_channels, outC=c.out_channels,
                                    image_size=c.image_size, downsample=c.downsample,
                                    return_2nd_last_layer_output=False)
        else:
            raise NotImplementedError(f"Input type {self.input_type} not supported")

    def forward_nhwc(self, x_nhwc):
        if self.input_type == "rgb":
            raise NotImplementedError("input_type='rgb' needs a 4-aligned channel pack; the shipped "
                                      "configs use 'rgbd'")
        return self.model.forward_nhwc(x_nhwc)

    def forward(self, img):
        return ops.nhwc_to_nchw(self.forward_nhwc(ops.nchw_to_nhwc(img.float())))


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/vision_encoder.py")
