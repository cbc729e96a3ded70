"""Mirror of the stage-1 training wrapper, reference creste/train_pefree.py:35-200
(`DistillationModel`), without the Lightning / Hydra / logging control plane (out of scope, SURVEY
section 8): model + LossManager + `training_step` / `validation_step` + `configure_optimizers`
(Adam(beta1, beta2, lr, eps) + ExponentialLR(gamma), train_pefree.py:176-200).

Lightning's automatic optimisation (zero_grad -> backward -> step around `training_step`) is folded
into `training_step` here.  Data parallelism follows the reference's DDP: one process per GPU,
per-rank batches and per-rank BatchNorm statistics, gradients averaged over ranks once per step --
ONE flat NCCL all-reduce of the 4.0 M-float gradient buffer (FlatAdam), then one fused Adam launch.
"""
import torch
from torch import nn

from creste_public_b200.config import as_cfg
from .models.distillation import DistillationBackbone  # noqa: F401  (resolved by name)
from .train_traversability import ExponentialLR, FlatAdam, broadcast_buffers
from .utils import loss_utils as lu
from .utils import train_utils as tu


class DistillationModel(nn.Module):
    """train_pefree.py:35-200.  `self.log` calls are collected in `self.logged`."""

    def __init__(self, model_cfg):
        super().__init__()
        model_cfg = as_cfg(model_cfg)
        self.model_cfg = model_cfg
        self.opt_cfg = model_cfg.optimizer
        self.lr_scheduler_cfg = model_cfg.lr_scheduler
        name = model_cfg.vision_backbone["class_name"]
        if name not in globals():
            raise NotImplementedError(f"Model {name} not found")
        self.model = globals()[name](model_cfg)
        self.log_keys = model_cfg.get("log_keys", [])
        self.loss = lu.LossManager(model_cfg)
        self.logged = {}
        self._opt = None
        self._sched = None

    def forward(self, x):
        return self.model(x)

    def configure_optimizers(self):
        if self.opt_cfg["name"] != "Adam":
            raise ValueError(f"Optimizer {self.opt_cfg['name']} not found.")
        self._opt = FlatAdam(self.model.parameters(), lr=self.opt_cfg["lr"],
                             betas=(self.opt_cfg["beta1"], self.opt_cfg["beta2"]))
        # the yaml's `eps: 1e-7` is never forwarded by the reference (train_pefree.py:177-181): torch's
        # default 1e-8 is what it trains with, and so does this mirror
        if self.lr_scheduler_cfg["name"] != "ExponentialLR":
            raise ValueError(f"LR scheduler {self.lr_scheduler_cfg['name']} not found.")
        self._sched = ExponentialLR(self._opt, self.lr_scheduler_cfg["gamma"])
        return [self._opt], [self._sched]

    def optimizers(self):
        if self._opt is None:
            self.configure_optimizers()
        return self._opt

    def _losses(self, inputs):
        if self.model_cfg.get("multiview_distillation", False):
            raise NotImplementedError("multiview_distillation is disabled in every shipped config")
        outputs = self(inputs["image"])
        with torch.no_grad():
            merged = tu.merge_dict(("inputs", inputs), ("outputs", outputs))
        loss_dict, meta = self.loss(merged)
        return outputs, loss_dict, meta, sum(w * v for w, v in loss_dict.values())

    def training_step(self, inputs):
        opt = self.optimizers()
        opt.zero_grad()
        broadcast_buffers(self.model, opt.group)
        _, loss_dict, meta, loss = self._losses(inputs)
        loss.backward()
        opt.step()
        self.logged.update({f"train/{k}": w * v.detach() for k, (w, v) in loss_dict.items()})
        self.logged.update({f"train/{k}": v.detach() for k, v in meta.items()})
        self.logged["train/loss"] = loss.detach()
        return {"loss": loss.detach()}

    def compute_gradient_norm(self):
        """train_pefree.py:110-118 as one reduction over the flat gradient buffer."""
        return self.optimizers().grad_norm()

    def validation_step(self, inputs):
        with torch.no_grad():
            _, loss_dict, meta, loss = self._losses(inputs)
        self.logged.update({f"val/{k}": w * v.detach() for k, (w, v) in loss_dict.items()})
        self.logged.update({f"val/{k}": v.detach() for k, v in meta.items()})
        self.logged["val/loss"] = loss.detach()
        return {"loss": loss}

    def on_train_epoch_end(self):
        if self._sched is not None:
            self._sched.step()
