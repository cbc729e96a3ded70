"""Mirror of reference creste/models/lfd.py (MaxEntIRL :20-392): frozen TerrainNet backbone +
VIN reward / value iteration + policy-propagation state-visitation frequencies."""
import os

import torch
from torch import nn

from creste_public_b200 import ops
from creste_public_b200.config import OmegaConf, open_dict
from creste_public_b200.engine import require_eval
from ..utils import train_utils as tu
from .blocks.vin import VIN  # noqa: F401  (globals() lookup)
from .terrainnet import TerrainNet  # noqa: F401


class MaxEntIRL(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.backbone_cfg = model_cfg.vision_backbone
        self.traversability_head_cfg = model_cfg.traversability_head
        self.policy_cfg = self.model_cfg.get("policy_kwargs", {})
        self.ckpt_path = self.model_cfg.get("ckpt_path", "")
        self.weights_path = self.model_cfg.get("weights_path", "")
        self.map_size = self.model_cfg.get("map_size", [64, 128])
        self.policy_method = self.model_cfg.get("policy_method", "fc")
        self.goal_cfg = self.model_cfg.get("goal_kwargs", {})
        self.action_horizon = self.model_cfg.get("action_horizon")
        self.solve_mdp = self.model_cfg.get("solve_mdp", False)
        self.zero_terminal_state = self.model_cfg.get("zero_terminal_state", False)
        self.register_buffer("dynamics", torch.tensor(
            [[-1, -1], [-1, 0], [-1, 1], [0, -1], [0, 1], [1, -1], [1, 0], [1, 1]], dtype=torch.long))
        fov = tu.create_trapezoidal_fov_mask(self.map_size[0] * 2, self.map_size[1], 70, 70, 0, 100)
        fov = fov.view(1, 1, self.map_size[0] * 2, self.map_size[1])
        self.fov_mask = fov[:, :, : self.map_size[0], : self.map_size[1]]   # plain attribute (ref)
        self._fov_dev = None
        n_actions = self.traversability_head_cfg["net_kwargs"]["qvalue_cfg"]["dims"][-1]
        if self.policy_method == "pp":
            # one-hot "where did the mass come from" kernels (reference lfd.py:58-70): action a
            # moves mass by dynamics[a], so the source of cell s is s - dynamics[a], i.e. tap
            # (1 - dy, 1 - dx) of a cross-correlation.  creste_svf has the same rule baked in.
            tp = torch.zeros(8, 1, 3, 3)
            for a in range(n_actions):
                dy, dx = self.dynamics[a].tolist()
                tp[a, 0, 1 - dy, 1 - dx] = 1.0
            self.register_buffer("transition_probs", tp)
        elif self.policy_method == "fc":
            raise NotImplementedError("policy_method='fc' (iterative_policy_rollout) is not used by "
                                      "the shipped configs (policy_method: 'pp')")
        else:
            raise ValueError(f"Policy method {self.policy_method} not found.")
        if "TerrainNet" not in self.backbone_cfg["project_name"]:
            raise ValueError(f"Model {self.backbone_cfg['project_name']} not found.")
        if self.backbone_cfg["load_setting"] == "strict_unfreezesplat":
            # the reference leaves cam2map trainable in this mode and optimises it in stage 3; here the
            # backbone runs without a graph inside MaxEntIRL.forward, so the splat layer would silently
            # never train -- refused until the stage-3 step differentiates the splat
            raise NotImplementedError("load_setting='strict_unfreezesplat' (trainable splat layer in stage 3) "
                                      "is not implemented; use 'strict_freeze' (the shipped stage-3 configs)")
        with open_dict(self.backbone_cfg):
            if self.backbone_cfg["load_setting"] != "strict_freeze":
                self.backbone_cfg["load_setting"] = "strict_freeze"
        self.backbone = TerrainNet(OmegaConf.create(self.backbone_cfg))
        if os.path.exists(self.backbone_cfg["weights_path"]):
            self.backbone.load_weights(self.backbone_cfg["weights_path"])
        self.traversability_head = globals()[self.traversability_head_cfg["value_iterator"]](
            **self.traversability_head_cfg["net_kwargs"])
        self.freeze_backbone = self.model_cfg.get("freeze_backbone", True)
        self.freeze_head = self.model_cfg.get("freeze_head", False)
        self.load_strict = self.model_cfg.get("load_strict", True)
        if os.path.isfile(self.weights_path) and not os.path.isfile(self.ckpt_path):
            self.load_weights(self.weights_path)

    def load_weights(self, weights_path):
        sd = torch.load(weights_path, weights_only=False)["state_dict"]
        sd = {k.replace("model.", "", 1): v for k, v in sd.items() if k.startswith("model.")}
        self.load_state_dict(sd, strict=self.load_strict)
        if self.freeze_backbone:
            self.backbone.eval()
            for p in self.backbone.parameters():
                p.requires_grad = False
        if self.freeze_head:
            self.traversability_head.eval()
            for p in self.traversability_head.parameters():
                p.requires_grad = False

    def _fov(self, device):
        if self._fov_dev is None or self._fov_dev.device != device:
            self._fov_dev = self.fov_mask[0, 0].to(torch.uint8).to(device).contiguous()
        return self._fov_dev

    def expected_state_visitation_frequency(self, policy, expert):
        """policy [B,8,H,W]; expert [B,T,3,3] SE(2) poses in un-pooled BEV cells ->
        exp_svf [B,H,W], state_preds [B,T,2] int64, state_preds_grid [B,H,W]
        (reference lfd.py:156-277; one kernel launch per call: creste_svf)."""
        B, A, H, W = policy.shape
        ds = self.traversability_head_cfg["net_kwargs"]["reward_cfg"]["ds"]
        rc = expert[:, :, :2, 2].to(policy.device).float().contiguous()
        method = self.policy_cfg["method"]
        if method not in ("sharpen", "none"):
            raise ValueError(f"Policy method {method} not found.")
        svf, states, grid = ops.svf(policy, rc, self._fov(policy.device), int(self.action_horizon),
                                    ds, method == "sharpen",
                                    float(self.policy_cfg.get("temperature", 1.0)),
                                    bool(self.zero_terminal_state))
        return {"exp_svf": svf, "state_preds_grid": grid, "state_preds": states}

    def forward(self, inputs):
        # The frozen backbone follows its own .training flag like the reference: eval = the fused inference
        # engine; train = batch-statistics BatchNorm + running-stat updates (what the reference does when no
        # stage-3 weights file froze it, lfd.py:141-145) -- never with a graph (frozen: no_grad).
        image, p2p = inputs[0], inputs[1]
        with torch.no_grad():
            outputs, preds_nhwc = self.backbone.forward_full((image, p2p))
        keys = self.traversability_head.reward_cfg.input_keys
        prefixes = [k[: -len("_preds")] for k in keys]
        preds = [preds_nhwc[p] for p in prefixes]
        if not self.solve_mdp:
            outputs.update(self.traversability_head.forward_nhwc(preds, None, False))
            return outputs
        assert len(inputs) > 2, "Goal location required for MDP solver"
        expert = inputs[2]
        B, _, H, W = outputs["bev_features"].shape
        map_ds = W // self.map_size[1]
        S = expert[:, :, :2, 2].long() // map_ds
        S[:, :, 0] = S[:, :, 0].clamp(0, self.map_size[0] - 1)
        S[:, :, 1] = S[:, :, 1].clamp(0, self.map_size[1] - 1)
        if "method" in self.goal_cfg:
            raise NotImplementedError("goal_kwargs is unused by the shipped configs")
        outputs.update(self.traversability_head.forward_nhwc(preds, S, solve_mdp=True))
        with torch.no_grad():
            outputs.update(self.expected_state_visitation_frequency(outputs["policy"], expert))
        return outputs


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/lfd.py")
