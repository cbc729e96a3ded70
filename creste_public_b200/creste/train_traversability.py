"""Mirror of the stage-3 training wrapper, reference creste/train_traversability.py:34-330
(`MaxEntIRLModel`), without the Lightning / Hydra / logging control plane (out of scope, SURVEY
section 8): model + LossManager + manual-optimisation `training_step` / `validation_step` +
`configure_optimizers` (Adam(beta1, beta2, lr) + ExponentialLR(gamma)).

Data parallelism follows the reference's DDP: one process per GPU, per-rank batches, gradients
averaged over ranks once per step.  Here that exchange is ONE flat NCCL all-reduce of the
102 866-float gradient buffer of the reward head (411 KB), followed by a fused Adam kernel
(`creste_adam_step`) on flat parameter / moment buffers.
"""
import torch
from torch import nn

from creste_public_b200 import engine, ops
from creste_public_b200.config import as_cfg
from .models.lfd import MaxEntIRL
from .utils import loss_utils as lu
from .utils import train_utils as tu


def _world(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def broadcast_buffers(module, group=None):
    """DDP's `broadcast_buffers=True` (the default Lightning's DDPStrategy keeps): before every forward
    rank 0's buffers -- the BatchNorm running statistics and step counters -- overwrite the other
    ranks'.  The batch statistics used for normalisation stay per-rank (no SyncBN), as in the reference.
    One coalesced broadcast of the floating-point buffers + one of the integer ones; no-op at world 1."""
    if _world(group) <= 1:
        return
    import torch.distributed as dist
    from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors
    src = dist.get_global_rank(group, 0) if group else 0
    bufs = [b for b in module.buffers() if b is not None and b.numel() > 0]
    for sel in (lambda b: b.is_floating_point(), lambda b: not b.is_floating_point()):
        by_dtype = {}
        for b in bufs:
            if sel(b):
                by_dtype.setdefault(b.dtype, []).append(b)
        for group_bufs in by_dtype.values():
            flat = _flatten_dense_tensors(group_bufs)
            dist.broadcast(flat, src=src, group=group)
            with torch.no_grad():
                for b, f in zip(group_bufs, _unflatten_dense_tensors(flat, group_bufs)):
                    b.copy_(f)


class FlatAdam:
    """torch.optim.Adam semantics (no amsgrad / weight decay) on ONE flat buffer: parameters and
    their .grad tensors are re-pointed to views of flat fp32 buffers, so the data-parallel
    exchange is a single all-reduce and the update a single kernel launch."""

    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        # every parameter starts on a 128-byte boundary of the flat buffers: kernels read parameters
        # (biases, BatchNorm vectors) with 16-byte vector loads straight from these views
        al = lambda k: (k + 31) // 32 * 32
        n = sum(al(p.numel()) for p in self.params)
        self.flat_p = torch.zeros(n, device=dev)
        self.flat_g = torch.zeros(n, device=dev)
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.views = []
        o = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + k].view_as(p)
                gv = self.flat_g[o:o + k].view_as(p)
                p.grad = gv
                self.views.append(gv)
                o += al(k)
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.steps = 0
        self.group = process_group
        # DistributedDataParallel broadcasts rank 0's parameters when it wraps the module (the reference
        # trains under Lightning's DDPStrategy): without it replicas initialised from per-rank RNG
        # streams would average gradients of DIFFERENT weights and drift apart silently
        if _world(self.group) > 1:
            import torch.distributed as dist
            dist.broadcast(self.flat_p, src=dist.get_global_rank(self.group, 0) if self.group else 0,
                           group=self.group)

    def zero_grad(self):
        self.flat_g.zero_()
        for p, gv in zip(self.params, self.views):
            p.grad = gv

    def _gather_grads(self):
        for p, gv in zip(self.params, self.views):
            if p.grad is None:
                continue
            if p.grad.data_ptr() != gv.data_ptr():     # autograd replaced the tensor: copy back
                gv.copy_(p.grad)
                p.grad = gv

    def step(self):
        import torch.distributed as dist
        self._gather_grads()
        scale = 1.0
        if _world(self.group) > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)
            scale = 1.0 / dist.get_world_size(self.group)
        self.steps += 1
        ops.adam_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0],
                      self.betas[1], self.eps, self.steps, scale)
        # the kernel wrote the parameters through the flat buffer's raw pointer: advance their version
        # counters so the eval-path pack caches (engine.PackCache) see the new values
        engine.mark_written(self.params)

    def grad_norm(self):
        """train_traversability.py:110-118 without the per-parameter .item() loop."""
        return torch.sqrt(ops.row_dot(self.flat_g.view(1, -1), self.flat_g.view(1, -1)))[0]


class ExponentialLR:
    def __init__(self, optimizer, gamma):
        self.opt, self.gamma = optimizer, float(gamma)

    def step(self):
        self.opt.lr *= self.gamma


class MaxEntIRLModel(nn.Module):
    """train_traversability.py:34-330.  `self.log` calls are collected in `self.logged`."""

    def __init__(self, model_cfg):
        super().__init__()
        model_cfg = as_cfg(model_cfg)
        self.model_cfg = model_cfg
        self.batch_size = int(model_cfg.get("batch_size", 1))
        self.opt_cfg = model_cfg.get("optimizer", {"name": "Adam", "beta1": 0.9, "beta2": 0.999,
                                                   "lr": 5e-4})
        self.lr_scheduler_cfg = model_cfg.get("lr_scheduler", {"name": "ExponentialLR", "gamma": 0.96})
        self.loss = lu.LossManager(model_cfg)
        self.model = MaxEntIRL(model_cfg)
        self.logged = {}
        self._opt = None
        self._sched = None

    def forward(self, x):
        return self.model(x)

    def configure_optimizers(self):
        if self.opt_cfg["name"] != "Adam":
            raise ValueError(f"Optimizer {self.opt_cfg['name']} not found.")
        for p in self.model.backbone.parameters():      # frozen backbone (lfd.py:141-145)
            p.requires_grad = False
        self.model.backbone.eval()
        self._opt = FlatAdam(self.model.traversability_head.parameters(), lr=self.opt_cfg["lr"],
                             betas=(self.opt_cfg["beta1"], self.opt_cfg["beta2"]))
        if self.lr_scheduler_cfg["name"] != "ExponentialLR":
            raise ValueError(f"LR scheduler {self.lr_scheduler_cfg['name']} not found.")
        self._sched = ExponentialLR(self._opt, self.lr_scheduler_cfg["gamma"])
        return [self._opt], [self._sched]

    def optimizers(self):
        if self._opt is None:
            self.configure_optimizers()
        return self._opt

    def _run(self, data, task):
        image, p2p = data["image"], data["p2p"]
        expert = data["traversability_label"]
        outputs = self.model((image, p2p, expert))
        with torch.no_grad():
            merged = tu.merge_dict(("inputs", data), ("outputs", outputs))
            merged["task"] = task
        return outputs, merged

    def training_step(self, inputs):
        batch, _, _ = inputs
        loss = 0.0
        for task, data in batch.items():
            opt = self.optimizers()
            opt.zero_grad()
            broadcast_buffers(self.model, opt.group)
            _, merged = self._run(data, task)
            loss_dict, meta = self.loss(merged)
            loss = loss + sum(w * v for w, v in loss_dict.values())
            loss.backward()
            opt.step()
            self.logged.update({f"train/{k}": w * v.detach() for k, (w, v) in loss_dict.items()})
            self.logged.update({f"train/{k}": v.detach() for k, v in meta.items()})
        self.logged["train/loss"] = loss.detach()
        return {"loss": loss}

    def validation_step(self, inputs):
        """Lightning runs validation under torch.no_grad(): the reward map carries no graph there, so the
        SMODICE gradient penalty is 0 and `val/loss` is the visitation term alone (loss_utils.py:1208)."""
        batch, _, _ = inputs
        loss = 0.0
        for task, data in batch.items():
            with torch.no_grad():
                _, merged = self._run(data, task)
                loss_dict, meta = self.loss(merged)
            loss = sum(w * v for w, v in loss_dict.values())
            self.logged.update({f"val/{k}": w * v.detach() for k, (w, v) in loss_dict.items()})
            self.logged.update({f"val/{k}": v.detach() for k, v in meta.items()})
        self.logged["val/loss"] = loss.detach() if torch.is_tensor(loss) else loss
        return {"loss": loss}

    def on_train_epoch_end(self):
        if self._sched is not None:
            self._sched.step()


class HeadStep:
    """Head-only stage-3 step (primary variant): the BEV head
    predictions are given, everything downstream runs -- max-pool/crop, reward FCN (train-mode
    BatchNorm, autograd), value iteration, state-visitation frequencies + rollout, MaxEntIRLLoss
    (with the double-backward gradient penalty), backward, gradient all-reduce, Adam."""

    def __init__(self, model: MaxEntIRL, loss_manager, lr=5e-4, betas=(0.9, 0.999)):
        self.model, self.loss = model, loss_manager
        self.opt = FlatAdam(model.traversability_head.parameters(), lr=lr, betas=betas)

    def __call__(self, feat_map, expert, fov_mask, counterfactuals):
        m = self.model
        self.opt.zero_grad()
        broadcast_buffers(m.traversability_head, self.opt.group)
        keys = m.traversability_head.reward_cfg.input_keys
        Wo = feat_map[keys[0]].shape[-1]
        map_ds = Wo // m.map_size[1]
        S = expert[:, :, :2, 2].long() // map_ds
        S[:, :, 0] = S[:, :, 0].clamp(0, m.map_size[0] - 1)
        S[:, :, 1] = S[:, :, 1].clamp(0, m.map_size[1] - 1)
        outputs = m.traversability_head(feat_map, S, solve_mdp=True)
        with torch.no_grad():
            outputs.update(m.expected_state_visitation_frequency(outputs["policy"], expert))
        td = {f"outputs/{k}": v for k, v in outputs.items()}
        td.update({"inputs/traversability_label": expert, "inputs/fov_mask": fov_mask,
                   "inputs/counterfactuals_label": counterfactuals, "task": None})
        loss_dict, meta = self.loss(td)
        loss = sum(w * v for w, v in loss_dict.values())
        loss.backward()
        self.opt.step()
        return loss.detach(), outputs, meta


class GraphedHeadStep(HeadStep):
    """HeadStep whose forward + loss + backward is replayed from a CUDA graph (static shapes).

    The eager step issues ~800 small launches (ours + ATen's) from Python for ~8 ms of GPU work at B = 8, 256 x 256:
    the host, not the GPU, bounds it.  The label-only half of the loss (MaxEntIRLLoss.prepare_labels: expert and
    counterfactual visitation rasters -- per-sample Python lists and one `.item()` each) runs eagerly BEFORE the
    replay; everything that depends on the model (reward FCN, value iteration, state-visitation frequencies, the loss
    with its double-backward gradient penalty, backward) is captured once; the gradient exchange and the fused Adam
    launch stay eager after it (no collective and no step-dependent scalar inside the graph).  Same arithmetic, same
    kernels, same order as HeadStep: the two agree bit for bit (tests/test_train_gpu.py)."""

    def __init__(self, model, loss_manager, example, lr=5e-4, betas=(0.9, 0.999), warmup=2):
        super().__init__(model, loss_manager, lr=lr, betas=betas)
        from creste_public_b200.creste.utils.loss_utils import MaxEntIRLLoss
        assert len(loss_manager.losses) == 1 and isinstance(loss_manager.losses[0], MaxEntIRLLoss), \
            "GraphedHeadStep captures the MaxEntIRLLoss-only configuration of train_traversability.py"
        self.irl = loss_manager.losses[0]
        feat_map, expert, fov_mask, cfs = example
        self.s_feat = {k: v.clone() for k, v in feat_map.items()}
        self.s_expert = expert.clone()
        head = model.traversability_head
        self.s_labels = {k: (v.clone() if torch.is_tensor(v) else v)
                         for k, v in self._labels(expert, fov_mask, cfs).items()}
        bufs = list(head.buffers())
        saved = [b.clone() for b in bufs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        from creste_public_b200 import _lib
        with torch.cuda.stream(side):
            for _ in range(warmup):
                n0 = _lib.lib().creste_launch_count()
                self._fwd_bwd()
                self.launches_per_replay = int(_lib.lib().creste_launch_count() - n0)   # our kernels inside the graph
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.s_loss, self.s_out, self.s_meta = self._fwd_bwd()
        with torch.no_grad():
            for b, s_ in zip(bufs, saved):
                b.copy_(s_)
        self.opt.zero_grad()

    def _labels(self, expert, fov_mask, cfs):
        m = self.model
        td = {"inputs/traversability_label": expert, "inputs/fov_mask": fov_mask, "inputs/counterfactuals_label": cfs}
        return self.irl.prepare_labels(td, (expert.shape[0], m.map_size[0], m.map_size[1]), device=expert.device)

    def _fwd_bwd(self):
        m = self.model
        self.opt.zero_grad()
        keys = m.traversability_head.reward_cfg.input_keys
        Wo = self.s_feat[keys[0]].shape[-1]
        map_ds = Wo // m.map_size[1]
        S = self.s_expert[:, :, :2, 2].long() // map_ds
        S[:, :, 0] = S[:, :, 0].clamp(0, m.map_size[0] - 1)
        S[:, :, 1] = S[:, :, 1].clamp(0, m.map_size[1] - 1)
        outputs = m.traversability_head(self.s_feat, S, solve_mdp=True)
        with torch.no_grad():
            outputs.update(m.expected_state_visitation_frequency(outputs["policy"], self.s_expert))
        td = {f"outputs/{k}": v for k, v in outputs.items()}
        assert self.irl.config.get("logvar_key", None) is None
        ld, md = self.irl.loss_from_labels(td, self.s_labels)
        name = self.irl.name
        # Loss.forward / LossManager.forward: (weight, value) pairs under "<loss name>/<key>"
        loss_dict = {f"{name}/{k}": (self.irl.weight * 1.0, v) for k, v in ld.items()}
        meta = {f"{name}/{k}": v for k, v in md.items()}
        loss = sum(w * v for w, v in loss_dict.values())
        loss.backward()
        return loss.detach(), outputs, meta

    def __call__(self, feat_map, expert, fov_mask, counterfactuals):
        labels = self._labels(expert, fov_mask, counterfactuals)         # eager: host lists, H2D copies, .item()
        with torch.no_grad():
            for k, v in feat_map.items():
                self.s_feat[k].copy_(v, non_blocking=True)
            self.s_expert.copy_(expert, non_blocking=True)
            for k, v in labels.items():
                if torch.is_tensor(v):
                    self.s_labels[k].copy_(v, non_blocking=True)
        broadcast_buffers(self.model.traversability_head, self.opt.group)
        self.graph.replay()
        self.opt.step()
        engine.mark_written(list(self.model.traversability_head.buffers()))
        return self.s_loss, self.s_out, self.s_meta
