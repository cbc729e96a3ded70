"""Mirror of reference creste/models/blocks/vin.py (VIN :21-155): reward FCN + value iteration.

value_iteration_manual is one persistent cooperative launch (creste_vi_solve) instead of ~690
x (8 launches + .item() sync)."""
import torch
from torch import nn

from creste_public_b200 import ops
from creste_public_b200.config import OmegaConf
from creste_public_b200.engine import require_eval
from .conv import MultiScaleFCN  # noqa: F401  (globals() lookup)


class VIN(nn.Module):
    def __init__(self, reward_cfg, qvalue_cfg):
        super().__init__()
        self.reward_cfg, self.qvalue_cfg = reward_cfg, qvalue_cfg
        self.discount = qvalue_cfg.get("discount", 0.95)
        self.r = globals()[self.reward_cfg["name"]](OmegaConf.create(self.reward_cfg["net_kwargs"]))
        assert len(self.qvalue_cfg.kernels) == 1, "Only single layer Q value network supported"
        # the fixed 0.1/0.8/0.1 transition stencil (reference vin.py:36-46); the CUDA kernel has
        # the same constants baked in -- the buffer exists for state_dict compatibility
        w = torch.zeros(self.qvalue_cfg.dims[1], 1, 3, 3)
        dest = [(0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2)]
        ring = [(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (1, 0)]  # clockwise
        for a, c in enumerate(dest[: self.qvalue_cfg.dims[1]]):
            i = ring.index(c)
            for pos, val in ((ring[(i - 1) % 8], 0.1), (c, 0.8), (ring[(i + 1) % 8], 0.1)):
                w[a, 0, pos[0], pos[1]] = val
        self.register_buffer("w", w)
        self.max_sweeps = 8192
        self.last_vi_info = None      # device int32[2]: (sweeps, hit_max) of the last solve

    def value_iteration_manual(self, r, goal, threshold=0.001, discount=0.95):
        """r [B,1,H,W] -> (v [B,1,H,W], policy [B,8,H,W], q [B,8,H,W]); `goal` is unused, as in
        the reference."""
        v, q, pi, info = ops.vi_solve(r, discount, threshold, self.max_sweeps)
        self.last_vi_info = info
        return v, pi, q

    def forward_nhwc(self, preds_nhwc, S, solve_mdp=False):
        """preds_nhwc: the head predictions in NHWC, ordered as reward_cfg.input_keys."""
        ds = self.reward_cfg.ds
        if ds != 2:
            raise NotImplementedError("reward_cfg.ds must be 2 (2x2 max-pool kernel)")
        N, Ho, Wo, _ = preds_nhwc[0].shape
        rows = (Ho // ds) // 2
        with torch.no_grad():
            iv_nhwc, iv_nchw = ops.maxpool2_concat([p.detach() for p in preds_nhwc], rows_out=rows,
                                                   want_nchw=True)
        train_graph = torch.is_grad_enabled() and any(p.requires_grad for p in self.r.parameters())
        if train_graph or self.r.training:
            # reference vin.py:116-119: input_view is a detached leaf that requires grad, so the
            # loss can take d(sum r)/d(input_view) (and differentiate it again w.r.t. the weights).
            # A train-mode head under no_grad (Lightning validation never does this, a user may) takes the
            # same path without a graph: BatchNorm batch statistics, as nn.Module.train() implies.
            from creste_public_b200 import autograd as ag
            if train_graph:
                iv_nchw.requires_grad_(True)
            r_nhwc = self.r.forward_autograd_nhwc(ag.ToNHWC.apply(iv_nchw))
            r = r_nhwc.view(N, 1, rows, Wo // ds)                   # C == 1: same memory as NCHW
            r_graph, r_nhwc = r, r_nhwc.detach()
        else:
            r_nhwc = self.r.forward_nhwc(iv_nhwc)                   # [N, rows, Wo/2, 1]
            r = r_nhwc.view(N, 1, rows, Wo // ds)
            r_graph = r
        # full-resolution copy: bilinear resize to (Ho//2, Wo) placed in the top half (:121-125)
        with torch.no_grad():
            full = torch.zeros(N, 1, Ho, Wo, device=r.device)
            up = ops.upsample_concat(None, _pad4(r_nhwc), (Ho // 2, Wo), None)   # C padded to 4
            full[:, 0, : Ho // 2, :] = up[..., 0]
        prefix = self.reward_cfg["output_prefix"][0]
        outputs = {prefix: r_graph, f"{prefix}_full": full, "input_view": iv_nchw}
        if not solve_mdp:
            return outputs
        assert S is not None, "No expert demonstrations given but solve mdp is True"
        with torch.no_grad():   # vin.py:134: no gradients through value iteration
            v, policy, q = self.value_iteration_manual(r_graph.detach(), S[:, -1, :], threshold=0.001,
                                                       discount=self.discount)
        outputs.update({"policy": policy, "q_estimate": q, "value_estimate": v})
        return outputs

    def forward(self, feat_map, S, solve_mdp=False):
        preds = [ops.nchw_to_nhwc(feat_map[k].float()) for k in self.reward_cfg.input_keys]
        return self.forward_nhwc(preds, S, solve_mdp)


def _pad4(x_nhwc1):
    """[N,H,W,1] -> [N,H,W,4] (zero channels) so the float4 upsample kernel can be reused."""
    N, H, W, _ = x_nhwc1.shape
    out = torch.zeros(N, H, W, 4, device=x_nhwc1.device)
    out[..., 0] = x_nhwc1[..., 0]
    return out


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/blocks/vin.py")
