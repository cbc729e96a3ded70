"""Hot-path part of reference creste/utils/depth_utils.py: convert_to_metric_depth_differentiable
(:300-313), convert_to_metric_depth (:316-343) and bin_depths (:346-383).  The softmax expectation and the
binning are one kernel launch each (creste_depth_expectation / creste_bin_depths); the image / point-cloud
I/O helpers of that file are dataset tooling and are not mirrored."""
import math

from creste_public_b200 import ops


def convert_to_metric_depth_differentiable(depth_logits, mode, depth_min, depth_max, num_bins):
    """depth_logits NCHW [B,D,H,W] -> expected depth [B,H,W] in the units of depth_min / depth_max (mm in the
    shipped configs).  Inside the training graphs the differentiable form is autograd.DepthExpectFn; this
    entry point is the value (inference) path."""
    if int(depth_logits.shape[1]) != int(num_bins):
        raise ValueError(f"{int(depth_logits.shape[1])} logit channels for num_bins={num_bins}")
    if depth_logits.requires_grad:
        from creste_public_b200 import autograd as ag
        return ag.depth_expectation_nchw(depth_logits, float(depth_min), float(depth_max), 1.0)
    m, _ = ops.depth_expectation(ops.nchw_to_nhwc(depth_logits.float()), float(depth_min), float(depth_max), 1.0)
    return m


def convert_to_metric_depth(depth_bin, mode, depth_min, depth_max, num_bins):
    """Bin index -> depth value (elementwise on whatever tensor the caller holds)."""
    if mode == "UD":
        return depth_bin * ((depth_max - depth_min) / num_bins) + depth_min
    if mode == "LID":
        bin_size = 2 * (depth_max - depth_min) / (num_bins * (1 + num_bins))
        return depth_min + 0.5 * bin_size * depth_bin * (depth_bin + 1)
    if mode == "SID":
        return (math.exp(math.log(1 + depth_max) - math.log(1 + depth_min)) * depth_bin / num_bins) + \
            math.log(1 + depth_min)
    raise NotImplementedError(mode)


def bin_depths(depth_map, mode, depth_min, depth_max, num_bins, target=False):
    """Depth map -> bin indices (float, or int64 with invalid -> num_bins when target=True)."""
    if mode not in ops.BIN_MODES:
        raise NotImplementedError(mode)
    return ops.bin_depths(depth_map, mode, depth_min, depth_max, num_bins, target)


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "utils/depth_utils.py")
