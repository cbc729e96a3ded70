"""Mirror of the stage-2 training wrapper, reference creste/train_ssc.py:43-269 (`TerrainNetModel`), without the
Lightning / Hydra / logging / visualisation control plane (out of scope): TerrainNet + LossManager +
`training_step` / `validation_step` + `configure_optimizers` (Adam(beta1, beta2, lr) over the parameters that
require grad + ExponentialLR(gamma)) + the backbone freeze / unfreeze schedule (:69-88).

Lightning's automatic optimisation (zero_grad -> backward -> step around `training_step`) is folded into
`training_step`.  Data parallelism follows the reference's DDP: per-rank batches and BatchNorm statistics, rank 0's
buffers broadcast before every forward, ONE flat NCCL all-reduce of the gradient buffer (FlatAdam) -- and, inside
SupPixelConLoss, the all-gather of the sampled pixel embeddings across ranks (the path's one other collective).
"""
import torch
from torch import nn

from creste_public_b200.config import as_cfg
from .models.terrainnet import TerrainNet
from .train_traversability import ExponentialLR, FlatAdam, broadcast_buffers
from .utils import loss_utils as lu
from .utils import train_utils as tu


class TerrainNetModel(nn.Module):
    """train_ssc.py:43-269.  `self.log` calls are collected in `self.logged`."""

    def __init__(self, model_cfg):
        super().__init__()
        model_cfg = as_cfg(model_cfg)
        self.model_cfg = model_cfg
        self.opt_cfg = model_cfg.optimizer
        self.lr_scheduler_cfg = model_cfg.lr_scheduler
        self.loss = lu.LossManager(model_cfg)
        self.model = TerrainNet(model_cfg)
        self.freeze_backbone_epochs = model_cfg.get("freeze_backbone_epochs", 0)
        self.backbone_frozen = False
        self.log_keys = model_cfg.get("log_keys", [])
        self.current_epoch = 0
        self.logged = {}
        self._opt = None
        self._sched = None

    # ---- backbone freeze schedule (train_ssc.py:69-88)
    def freeze_backbone(self):
        for p in self.model.depthcomp.parameters():
            p.requires_grad = False
        self.backbone_frozen = True
        self._opt = None                  # the trainable set changed: rebuild the flat optimiser state lazily

    def unfreeze_backbone(self):
        self.model.depthcomp.unfreeze_backbone()
        self.backbone_frozen = False
        self._opt = None

    def on_train_epoch_start(self):
        if self.current_epoch >= self.freeze_backbone_epochs and self.backbone_frozen:
            self.unfreeze_backbone()
        elif self.current_epoch < self.freeze_backbone_epochs and not self.backbone_frozen:
            self.freeze_backbone()

    def on_train_epoch_end(self):
        if self._sched is not None:
            self._sched.step()
        self.current_epoch += 1

    def forward(self, x):
        return self.model(x)

    def configure_optimizers(self):
        if self.opt_cfg["name"] != "Adam":
            raise NotImplementedError(self.opt_cfg["name"])
        lr = self._opt.lr if self._opt is not None else self.opt_cfg["lr"]
        self._opt = FlatAdam((p for p in self.model.parameters() if p.requires_grad), lr=lr,
                             betas=(self.opt_cfg["beta1"], self.opt_cfg["beta2"]))
        if self.lr_scheduler_cfg["name"] != "ExponentialLR":
            raise NotImplementedError(self.lr_scheduler_cfg["name"])
        self._sched = ExponentialLR(self._opt, self.lr_scheduler_cfg["gamma"])
        return [self._opt], [self._sched]

    def optimizers(self):
        if self._opt is None:
            self.configure_optimizers()
        return self._opt

    def _losses(self, batch):
        loss, loss_dict_full, meta_full = 0.0, {}, {}
        for task, data in batch.items():
            outputs = self.model((data["image"], data["p2p"], data.get("immovable_depth_label", None)))
            with torch.no_grad():
                merged = tu.merge_dict(("inputs", data), ("outputs", outputs))
            merged["task"] = task
            loss_dict, meta = self.loss(merged)
            meta_full = tu.merge_loss_dict(meta_full, meta)
            loss_dict_full = tu.merge_loss_dict(loss_dict_full, loss_dict)
            loss = loss + sum(w * v for w, v in loss_dict.values())
        return loss, loss_dict_full, meta_full

    def training_step(self, inputs):
        batch, _, _ = inputs
        opt = self.optimizers()
        opt.zero_grad()
        broadcast_buffers(self.model, opt.group)
        loss, loss_dict, meta = self._losses(batch)
        loss.backward()
        opt.step()
        self.logged.update({f"train/{k}": w * v.detach() for k, (w, v) in loss_dict.items()})
        self.logged.update({f"train/{k}": v.detach() for k, v in meta.items()})
        self.logged["train/loss"] = loss.detach()
        return {"loss": loss.detach()}

    def validation_step(self, inputs):
        batch, _, _ = inputs
        with torch.no_grad():
            loss, loss_dict, meta = self._losses(batch)
        self.logged.update({f"val/{k}": w * v.detach() for k, (w, v) in loss_dict.items()})
        self.logged.update({f"val/{k}": v.detach() for k, v in meta.items()})
        self.logged["val/loss"] = loss.detach()
        return {"loss": loss}
