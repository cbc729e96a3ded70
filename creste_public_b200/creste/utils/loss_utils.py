"""Mirror of the stage-3 part of reference creste/utils/loss_utils.py: Loss (:25-59),
LossManager (:63-92) and MaxEntIRLLoss (:971-1259).

The per-sample reductions, normalisations, the visitation rasters and the SMODICE gradient
penalty run on the sm_100a kernels (creste_row_*, creste_expert_visitation,
creste_grad_penalty); the differentiable pieces are creste_public_b200.autograd Functions, so
`loss.backward()` reaches the reward-FCN weights through the first- and second-order graph
exactly as in the reference.  The stage-1 losses (CrossEntropyDepth, SmoothL1Depth, MSELoss) are one
fused kernel pass for the values and creste_ce_depth_bwd / creste_masked_mse_bwd for the gradients
(the Smooth-L1 term reads int64 bins and has none, as in the reference).  The stage-2 losses of
configs/model/ssc_sam/*.yaml -- SupPixelConLoss (:203-286), CrossEntropy (:379-474), SmoothL1Depth on the
differentiable soft-argmax depth, SmoothL1 (:576-604) -- are fused value + gradient kernels too (csrc/losses2.cu);
the data-dependent pixel selection of SupPixelConLoss (valid-mask gather, per-class random subsampling) is index
bookkeeping on the device, as in the reference.
"""
import weakref

import numpy as np
import torch
from torch import nn

from creste_public_b200 import autograd as ag
from creste_public_b200 import ops
from ..models.losses.supcon_loss import MultiPosConLoss
from . import train_utils as tu
from . import utils


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
            raise NotImplementedError(f"loss {config['name']} is outside the hot path (SURVEY.md section 8)")
        return globals()[config["name"]](config)


class _Stage1DepthValues:
    """CrossEntropyDepth and SmoothL1Depth share one kernel pass; the result is cached per
    (logits, label) pair so that the two Loss objects of the shipped config cost one launch."""
    _key, _val, _ref = None, None, None

    @classmethod
    def get(cls, tensor_dict, discretize, beta):
        logits = tensor_dict["outputs/depth_preds_logits"]
        bins = tensor_dict.get("outputs/depth_preds_bins")
        if bins is None:                       # a caller that hands over the logits alone: the bins ARE their arg-max
            bins = logits.detach().argmax(dim=1)
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
    """Reference loss_utils.py:530-573.  Stage 1 wires pred_key to depth_preds_bins (int64 class indices compared
    with metres: a value without a gradient, exactly as in the reference); stage 2 wires it to the soft-argmax
    depth_preds_metric, which IS differentiated (autograd.SmoothL1Fn over the valid LiDAR bins)."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.beta = config["beta"]

    def loss(self, tensor_dict):
        key = self.config["pred_key"]
        if key == "outputs/depth_preds_bins":
            acc = _Stage1DepthValues.get(tensor_dict, self.config["discretize"], self.beta)
            return {"depth/reg_loss": (acc[3] / acc[1]).float()}, {}
        pred = tensor_dict[key]                                         # [B*S, H, W] metres
        gt = tensor_dict[self.config["lab_key"]]
        B, S, H, W = gt.shape
        if pred.shape[0] != B * S or tuple(pred.shape[-2:]) != (H, W):
            raise NotImplementedError("multi-frame / resized depth labels are outside the hot path")
        disc = self.config["discretize"]
        gt = gt.reshape(B * S, H, W).to(pred.device).float().contiguous()
        with torch.no_grad():
            valid = ops.bin_depths(gt, disc["mode"], disc["depth_min"], disc["depth_max"], disc["num_bins"],
                                   target=True) != int(disc["num_bins"])
        if pred.requires_grad and torch.is_grad_enabled():
            return {"depth/reg_loss": ag.SmoothL1Fn.apply(pred.contiguous(), gt, valid, 1e-3, float(self.beta))}, {}
        acc = ops.smooth_l1(pred, gt, valid, 1e-3, float(self.beta))
        return {"depth/reg_loss": (acc[0] / acc[1]).float()}, {}


class SmoothL1(Loss):
    """Reference loss_utils.py:576-604 (elevation regression): channel 1 of the target becomes relative to channel
    0 IN PLACE unless `absolute` (the reference mutates its input the same way), non-finite targets are skipped."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.beta = config["beta"]
        self.pred_key, self.lab_key = config["pred_key"], config["lab_key"]
        self.absolute = config.get("absolute", False)
        if config.get("take_grad", False):
            raise NotImplementedError("SmoothL1(take_grad=True) is unused by the shipped configs")

    def loss(self, tensor_dict):
        pred, gt = tensor_dict[self.pred_key], tensor_dict[self.lab_key]
        if not self.absolute:
            gt[:, 1, :, :] = gt[:, 1, :, :] - gt[:, 0, :, :]
        g = gt.to(pred.device).float().contiguous()
        if pred.requires_grad and torch.is_grad_enabled():
            return {"val": ag.SmoothL1Fn.apply(pred.contiguous(), g, None, 1.0, float(self.beta))}, {}
        acc = ops.smooth_l1(pred, g, None, 1.0, float(self.beta))
        return {"val": (acc[0] / acc[1]).float()}, {}


def _class_weights(config, eps=1e-5):
    """1 / log(frequency + eps) from the text file named by config['class_weights'] (loss_utils.py:386-392)."""
    freq = np.loadtxt(config["class_weights"])
    return torch.from_numpy(1 / np.log(freq + eps)).float()


class CrossEntropy(Loss):
    """Reference loss_utils.py:379-474: class-weighted cross-entropy of the BEV logits over the FOV-masked cells
    (+ the accuracy meta value over the cells whose label is not 0); one fused kernel pass each way."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.num_class = config["num_class"]
        self.epsilon_w = 1e-5
        if "class_weights" in config:
            self.register_buffer("class_weights", _class_weights(config, self.epsilon_w))
            assert self.num_class == len(self.class_weights)
        else:
            self.class_weights = None
        self.mask_key = config.get("mask_key", "inputs/fov_mask")
        self.pred_key = config.get("pred_key", "outputs/inpainting_preds")
        self.lab_key = config.get("lab_key", "inputs/sem_label")
        self.ignore_index = config.get("ignore_index", None)
        self.task = config.get("task", "3d_ssc")
        self.class_dim = config.get("class_dim", -1)

    def loss(self, tensor_dict):
        pred = tensor_dict[self.pred_key]                               # [B,C,H,W]
        gt = tensor_dict[self.lab_key]
        with torch.no_grad():
            if self.class_dim < 0:
                gt_mode = torch.argmax(gt / (gt.sum(dim=1, keepdim=True) + self.epsilon_w), dim=1)
            else:
                gt_mode = gt[:, self.class_dim, :, :].long()
            gt_mode = gt_mode.to(pred.device).contiguous()
            mask = tensor_dict[self.mask_key].to(pred.device).to(torch.uint8).contiguous()
        ign = -100 if self.ignore_index is None else int(self.ignore_index)
        cw = None if self.class_weights is None else self.class_weights.to(pred.device)
        acc = ops.ce_weighted(pred.detach(), gt_mode, mask, cw, ign)
        if pred.requires_grad and torch.is_grad_enabled():
            val = ag.WeightedCEFn.apply(pred.contiguous(), gt_mode, mask, cw, ign, acc)
        else:
            val = (acc[0] / acc[1]).float()
        return {f"{self.task}/cls_loss": val}, {f"{self.task}/mIoU": (acc[2] / (acc[3] + self.epsilon_w)).float()}


class SupPixelConLoss(Loss):
    """Reference loss_utils.py:203-286: supervised pixel-contrastive loss on the BEV embedding head.  Labels are
    made unique across the batch (SAM masks are per-frame ids), the valid (labelled & in-FOV) cells are gathered,
    every class is randomly subsampled to the median class size (at most 1000), and the multi-positive contrastive
    loss is evaluated over those embeddings (all-gathered across ranks)."""

    def __init__(self, config):
        super().__init__(config.name if hasattr(config, "name") else config["name"], config)
        self.views = config.get("views", 1)
        self.temperature = config.get("temperature", 0.1)
        self.epsilon_w = 1e-5
        if "class_weights" in config:
            self.register_buffer("class_weights", _class_weights(config, self.epsilon_w))
            self.num_class = config["num_class"]
            assert self.num_class == len(self.class_weights)
        else:
            self.class_weights = None
        self.supcon_loss = MultiPosConLoss(temperature=self.temperature, class_weights=self.class_weights)
        self.ignore_index = config.get("ignore_index", -1)
        self.mask_key = config.get("mask_key", "inputs/fov_mask")
        self.pred_key = config.get("pred_key", "outputs/inpainting_preds")
        self.lab_key = config.get("lab_key", "inputs/sem_label")
        self.lab_suffix_key = self.lab_key.split("/")[-1]
        self.task = config.get("task", "3d_ssc")
        if self.views != 1:
            raise NotImplementedError("views != 1 is unused by the shipped configs")

    def loss(self, tensor_dict):
        preds = tensor_dict[self.pred_key]                              # [B,Z,H,W]
        gt_prob = tensor_dict[self.lab_key]
        B, Z, H, W = preds.shape
        with torch.no_grad():
            fov = tensor_dict[self.mask_key].to(preds.device)
            gt_label = (torch.argmax(gt_prob, dim=1) if gt_prob.shape[1] > 1 else gt_prob.squeeze(1)).to(preds.device)
            if self.lab_key == "inputs/3d_sam_label":
                gt_label = utils.remap_labels_in_batch(gt_label, ignore_idx=0)
            valid = ((gt_label != self.ignore_index) & fov).view(B, H, W)
            lab = gt_label.view(B, H, W)[valid]                         # [N]
            counts = torch.bincount(lab)
            nz = counts[counts.nonzero(as_tuple=True)].float()
            median_count = min(nz.median().int(), 1000)
            sel = tu.extract_max_per_class(lab, median_count, return_indices=True)
            lab = lab[sel]
            # flat cell index of every selected embedding in the [B,H,W] grid
            cells = valid.reshape(-1).nonzero(as_tuple=False).reshape(-1)[sel]
        # gather of the selected pixel embeddings: NCHW -> [N,Z] (torch index_select: differentiable plumbing)
        flat = preds.permute(0, 2, 3, 1).reshape(B * H * W, Z)
        feats = flat.index_select(0, cells)
        out = self.supcon_loss({"feats": feats, "labels": lab})
        base = f"{self.task}/{self.lab_suffix_key}/supcon"
        return {f"{base}/sem_loss": out["loss"], f"{base}/img_loss": out["image_loss"]}, {}


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

    def prepare_labels(self, tensor_dict, shape, device=None):
        """The label-only half of `loss` (reference loss_utils.py:1118-1180): FOV mask crop, expert visitation
        distribution, counterfactual visitation distributions.  Depends on the batch, not on the model, and contains
        every host-side step of the loss (the per-sample trajectory lists, the `.item()` of the interpolation length),
        so a caller may run it ahead of -- and outside -- a captured forward / backward graph.
        -> dict(mask uint8 [B,H,W] | None, svf [B,H,W], cf [B,H,W], has_cf float [B,1,1])"""
        gt = tensor_dict[self.lab_key]
        fov_mask = tensor_dict[self.fov_key]
        B, H, W = shape
        dev = device if device is not None else tensor_dict["outputs/traversability_preds"].device
        _, Ho, Wo = fov_mask.shape
        fov_mask = tu.resize_and_crop(fov_mask.to(dev).unsqueeze(1).byte(), (Ho // 2, Wo // 2),
                                      (0, H, 0, W)).squeeze(1).contiguous()        # uint8 [B,H,W]
        mask = fov_mask if self.use_fov_mask else None
        with torch.no_grad():
            _, svf = self.compute_expert_visitation(gt.to(dev), self.map_ds, self.map_sz)
            svf = ops.row_normalize(svf, mask, 1e-5)
            cf_svf_total = torch.zeros_like(svf)
            has_cf = torch.zeros(B, 1, 1, device=dev)
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
                    cf_svf_total[idx] = ops.row_normalize(cf.view(1, H, W), None, 1e-5)[0]
                    has_cf[idx] = 1.0
        return {"mask": mask, "svf": svf, "cf": cf_svf_total, "has_cf": has_cf}

    def loss_from_labels(self, tensor_dict, labels):
        """The model-dependent half of `loss`: device work only, static shapes (capturable in a CUDA graph)."""
        exp_svf = tensor_dict[self.pred_key]
        reward_preds = tensor_dict["outputs/traversability_preds"]
        state_features = tensor_dict["outputs/input_view"]
        reward_preds = reward_preds.squeeze(1)                         # [B,H,W]
        dev = reward_preds.device
        mask, svf, cf_svf_total, has_cf = labels["mask"], labels["svf"], labels["cf"], labels["has_cf"]
        with torch.no_grad():
            exp_svf = ops.row_normalize(exp_svf.detach().float(), mask, 1e-5)
            exp_svf_total = exp_svf.clone()
            if self.cf_key is not None and self.alpha is not None:
                # exp_svf[idx] = alpha * cf + (1 - alpha) * exp_svf[idx] for the samples that have counterfactuals
                exp_svf = torch.where(has_cf > 0, self.alpha * cf_svf_total + (1 - self.alpha) * exp_svf, exp_svf)

        # differentiable part: sums of rewards under the two visitation distributions; the
        # reference masks the reward (reward_preds * ones_mask) -- here the mask rides in RowDot
        svf_rewards = ag.RowDotFn.apply(reward_preds, svf, mask)
        exp_svf_rewards = ag.RowDotFn.apply(reward_preds, exp_svf, mask)
        mean_exp_svf_rewards = exp_svf_rewards.mean()
        mean_svf_rewards = svf_rewards.mean()
        visitation_loss = mean_exp_svf_rewards - mean_svf_rewards

        reward_penalty = torch.zeros((), device=dev)
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
            valid = (cf_rewards != 0).to(cf_rewards.dtype)
            cf_rewards = (cf_rewards * valid).sum()         # sum over the samples with a counterfactual reward
            opt_rewards = (opt_rewards * valid).sum()
        meta = {
            "reward_penalty": self.reward_weight * reward_penalty,
            "mean_expected_svf_rewards": mean_exp_svf_rewards,
            "mean_svf_rewards": mean_svf_rewards,
            "sum_cf_rewards": cf_rewards,
            "sum_opt_rewards": opt_rewards,
        }
        return {"maxentirl_loss": loss}, meta

    def loss(self, tensor_dict):
        B, H, W = tensor_dict[self.pred_key].shape
        return self.loss_from_labels(tensor_dict, self.prepare_labels(tensor_dict, (B, H, W)))


def _sum_rows(x):
    """[N,H,W] -> [H,W] sum over N (N is a handful of counterfactual trajectories; a torch
    reduction over <= a few hundred KB, as in the reference)."""
    return x.sum(dim=0)


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "utils/loss_utils.py")
