"""Mirror of the stage-3 part of reference creste/utils/loss_utils.py: Loss (:25-59),
LossManager (:63-92) and MaxEntIRLLoss (:971-1259).

The per-sample reductions, normalisations, the visitation rasters and the SMODICE gradient
penalty run on the sm_100a kernels (creste_row_*, creste_expert_visitation,
creste_grad_penalty); the differentiable pieces are creste_public_b200.autograd Functions, so
`loss.backward()` reaches the reward-FCN weights through the first- and second-order graph
exactly as in the reference.  The stage-1 losses (CrossEntropyDepth, SmoothL1Depth, MSELoss) are one
fused kernel pass for the values and creste_ce_depth_bwd / creste_masked_mse_bwd for the gradients
(the Smooth-L1 term reads int64 bins and has none, as in the reference); stage-2 losses are not mirrored.
"""
import weakref

import numpy as np
import torch
from torch import nn

from creste_public_b200 import autograd as ag
from creste_public_b200 import ops
from . import train_utils as tu


class Loss(nn.Module):
    def __init__(self, name, config):
        super().__init__()
        self.config = config
        self._name = name + config.get("tag", "")
        self.weight = config.get("weight", 1.0)
        self.task = config.get("task", None)

    def forward(self, tensor_dict):
        loss_dict, meta_data = self.loss(tensor_dict)
        ret = {}
        logvar_key = self.config.get("logvar_key", None)
        if logvar_key is not None:
            log_var = tensor_dict[logvar_key]
            w = 1.0 / (2.0 * torch.exp(log_var))
            ret["log_std"] = (1.0, 0.5 * log_var)
        else:
            w = 1.0
        ret.update({k: (self.weight * w, v) for k, v in loss_dict.items()})
        return ret, meta_data

    def loss(self, tensor_dict):
        raise Exception("Not Implemented!")

    @property
    def name(self):
        return self._name


class LossManager(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.losses = nn.ModuleList()
        for lc in config.loss:
            self.losses.append(self.get_loss(lc))

    def forward(self, tensor_dict):
        loss_dict, meta_data = {}, {}
        for loss in self.losses:
            if loss.task is None or loss.task == tensor_dict["task"]:
                ld, md = loss(tensor_dict)
                meta_data.update({f"{loss.name}/{k}": v for k, v in md.items()})
                loss_dict.update({f"{loss.name}/{k}": v for k, v in ld.items()})
        return loss_dict, meta_data

    def get_loss(self, config):
        if config["name"] not in globals():
            raise NotImplementedError(f"loss {config['name']} is outside the stage-3 hot path")
        return globals()[config["name"]](config)


class _Stage1DepthValues:
    """CrossEntropyDepth and SmoothL1Depth share one kernel pass; the result is cached per
    (logits, label) pair so that the two Loss objects of the shipped config cost one launch."""
    _key, _val, _ref = None, None, None

    @classmethod
    def get(cls, tensor_dict, discretize, beta):
        logits = tensor_dict["outputs/depth_preds_logits"]
        bins = tensor_dict["outputs/depth_preds_bins"]
        label = tensor_dict["inputs/depth_label"]
        key = (logits.data_ptr(), label.data_ptr(), logits._version, float(beta))
        # the cache is tied to the logits OBJECT (weak reference): a later step's logits may reuse the
        # address of a freed tensor with the same version counter
        if cls._key != key or cls._ref is None or cls._ref() is not logits:
            if discretize["mode"] != "UD":
                raise NotImplementedError("bin_depths modes other than 'UD' are unused by the shipped configs")
            if int(logits.shape[1]) != int(discretize["num_bins"]):
                raise ValueError(f"depth head emits {int(logits.shape[1])} bins, discretize.num_bins is "
                                 f"{discretize['num_bins']}")
            B, S, H, W = label.shape
            if logits.shape[0] != B * S or tuple(logits.shape[-2:]) != (H, W):
                raise NotImplementedError("multi-frame / resized depth labels are outside the hot path")
            cls._val = ops.stage1_depth_losses(logits, bins.reshape(B * S, H * W),
                                               label.reshape(B * S, H * W).to(logits.device),
                                               discretize["depth_min"], discretize["depth_max"], beta)
            cls._key = key
            cls._ref = weakref.ref(logits)
        return cls._val


class CrossEntropyDepth(Loss):
    """Reference loss_utils.py:477-527: mean cross-entropy of the depth logits against the UD-binned
    LiDAR depth over the valid pixels (+ the depth/acc meta value); one fused kernel pass for the
    value, creste_ce_depth_bwd for the gradient (autograd.CEDepthFn) when the logits carry one."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)

    def loss(self, tensor_dict):
        disc = self.config["discretize"]
        acc = _Stage1DepthValues.get(tensor_dict, disc, 0.5)
        logits = tensor_dict["outputs/depth_preds_logits"]
        if logits.requires_grad and torch.is_grad_enabled():
            label = tensor_dict["inputs/depth_label"]
            label = label.reshape(logits.shape[0], -1).to(logits.device).float().contiguous()
            val = ag.CEDepthFn.apply(logits, label, acc, float(disc["depth_min"]), float(disc["depth_max"]))
        else:
            val = (acc[0] / acc[1]).float()
        return {"depth/cls_loss": val}, {"depth/acc": (acc[2] / acc[1]).float()}


class SmoothL1Depth(Loss):
    """Reference loss_utils.py:530-573; pred_key is depth_preds_bins in the shipped config (int64
    class indices compared with metres), so -- exactly as in the reference -- this term has a value
    but contributes no gradient."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.beta = config["beta"]

    def loss(self, tensor_dict):
        if self.config["pred_key"] != "outputs/depth_preds_bins":
            raise NotImplementedError("SmoothL1Depth is wired to depth_preds_bins in the shipped configs")
        acc = _Stage1DepthValues.get(tensor_dict, self.config["discretize"], self.beta)
        return {"depth/reg_loss": (acc[3] / acc[1]).float()}, {}


class MSELoss(Loss):
    """Reference loss_utils.py:606-647 (overlap_only = False); differentiable via MaskedMSEFn."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.pred_key, self.lab_key = config["pred_key"], config["lab_key"]
        if config.get("overlap_only", False):
            raise NotImplementedError("MSELoss(overlap_only=True) is unused by the shipped configs")

    def loss(self, tensor_dict):
        pred, gt = tensor_dict[self.pred_key], tensor_dict[self.lab_key]
        assert pred.shape == gt.shape, (pred.shape, gt.shape)
        gt = gt.to(pred.device).float().contiguous()
        if pred.requires_grad and torch.is_grad_enabled():
            return {"loss": ag.MaskedMSEFn.apply(pred.contiguous(), gt)}, {}
        acc = ops.masked_mse(pred, gt)
        return {"loss": (acc[0] / acc[1]).float()}, {}


class MaxEntIRLLoss(Loss):
    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.pred_key = config["pred_key"]
        self.lab_key = config["lab_key"]
        self.fov_key = config["fov_key"]
        self.map_ds = config.get("map_ds", 2)
        self.map_sz = config.get("map_sz", [64, 128])
        self.maxent_weight = config.get("maxent_weight", 1.0)
        self.reward_weight = config.get("reward_weight", 0.1)
        self.use_fov_mask = config.get("use_fov_mask", False)
        self.alpha = config.get("alpha", None)
        self.cf_key = config.get("cf_key", None)

    @staticmethod
    def compute_expert_visitation(gt, map_ds, map_sz):
        """gt [B,T,3,3] poses or [B,T,2] (row, col), un-pooled BEV cells ->
        (None, visit_counts [B,H,W] in {0,1}).  Reference loss_utils.py:1055-1116 (the second,
        effective definition); the interpolated point list is not materialised."""
        xy = gt if gt.ndim == 3 else gt[:, :, :2, 2]
        xy = xy.contiguous()
        H, W = map_sz
        seg = (xy[:, 1:] - xy[:, :-1]) / map_ds
        # the reference's one host sync of this function: ceil(max segment length).item()
        max_steps = int(torch.ceil(torch.norm(seg, dim=-1)).long().max().item())
        return None, ops.expert_visitation(xy, map_ds, max_steps, H, W)

    def loss(self, tensor_dict):
        exp_svf = tensor_dict[self.pred_key]
        gt = tensor_dict[self.lab_key]
        fov_mask = tensor_dict[self.fov_key]
        reward_preds = tensor_dict["outputs/traversability_preds"]
        state_features = tensor_dict["outputs/input_view"]
        reward_preds = reward_preds.squeeze(1)                         # [B,H,W]
        dev = reward_preds.device
        _, Ho, Wo = fov_mask.shape
        B, H, W = exp_svf.shape
        fov_mask = tu.resize_and_crop(fov_mask.to(dev).unsqueeze(1).byte(), (Ho // 2, Wo // 2),
                                      (0, H, 0, W)).squeeze(1).contiguous()        # uint8 [B,H,W]
        mask = fov_mask if self.use_fov_mask else None

        with torch.no_grad():
            _, svf = self.compute_expert_visitation(gt.to(dev), self.map_ds, self.map_sz)
            svf = ops.row_normalize(svf, mask, 1e-5)
            exp_svf = ops.row_normalize(exp_svf.detach().float(), mask, 1e-5)
            cf_svf_total = torch.zeros_like(svf)
            exp_svf_total = exp_svf.clone()
            if self.cf_key is not None and self.alpha is not None:
                for idx, cf_dict in enumerate(tensor_dict[self.cf_key]):
                    if cf_dict is None:
                        continue
                    invalid = cf_dict["trajectories"][cf_dict["rank"] > 0]
                    if invalid.shape[0] == 0:
                        continue
                    invalid = torch.from_numpy(np.ascontiguousarray(invalid)).to(dev)
                    _, cf = self.compute_expert_visitation(invalid, self.map_ds, self.map_sz)
                    # sum over the trajectories, then normalise to a distribution over cells
                    cf = _sum_rows(cf)
                    cf = ops.row_normalize(cf.view(1, H, W), None, 1e-5)[0]
                    exp_svf[idx] = self.alpha * cf + (1 - self.alpha) * exp_svf[idx]
                    cf_svf_total[idx] = cf

        # differentiable part: sums of rewards under the two visitation distributions; the
        # reference masks the reward (reward_preds * ones_mask) -- here the mask rides in RowDot
        svf_rewards = ag.RowDotFn.apply(reward_preds, svf, mask)
        exp_svf_rewards = ag.RowDotFn.apply(reward_preds, exp_svf, mask)
        mean_exp_svf_rewards = exp_svf_rewards.mean()
        mean_svf_rewards = svf_rewards.mean()
        visitation_loss = mean_exp_svf_rewards - mean_svf_rewards

        reward_penalty = torch.tensor(0.0, device=dev)
        if reward_preds.requires_grad and self.reward_weight > 0:
            # sum of the (FOV-masked) reward map, as reward_preds.sum() in the reference
            rp_sum = ag.RowDotFn.apply(reward_preds, torch.ones_like(reward_preds), mask).sum()
            reward_grad = torch.autograd.grad(outputs=rp_sum, inputs=state_features,
                                              create_graph=True, retain_graph=True,
                                              only_inputs=True)[0]
            reward_penalty = ag.GradPenaltyFn.apply(reward_grad)

        loss = self.maxent_weight * visitation_loss + self.reward_weight * reward_penalty

        with torch.no_grad():
            r = reward_preds.detach()
            cf_rewards = ops.row_dot(r, cf_svf_total, mask)
            opt_rewards = ops.row_dot(r, exp_svf_total, mask)
            valid = cf_rewards != 0
            cf_rewards = cf_rewards[valid].sum()
            opt_rewards = opt_rewards[valid].sum()
        meta = {
            "reward_penalty": self.reward_weight * reward_penalty,
            "mean_expected_svf_rewards": mean_exp_svf_rewards,
            "mean_svf_rewards": mean_svf_rewards,
            "sum_cf_rewards": cf_rewards,
            "sum_opt_rewards": opt_rewards,
        }
        return {"maxentirl_loss": loss}, meta


def _sum_rows(x):
    """[N,H,W] -> [H,W] sum over N (N is a handful of counterfactual trajectories; a torch
    reduction over <= a few hundred KB, as in the reference)."""
    return x.sum(dim=0)
