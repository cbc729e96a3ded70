"""Mirror of reference creste/models/losses/supcon_loss.py: MultiPosConLoss (:56-115), the multi-positive
contrastive loss of https://arxiv.org/pdf/2306.00984.pdf that SupPixelConLoss applies to the BEV pixel embeddings.

    feats -> F.normalize -> all-gather over the data-parallel ranks (WITH gradient: torch.distributed.nn.all_gather)
    -> logits = feats @ all_feats.T / T with the self column masked -> cross-entropy against the uniform
    distribution over the same-label columns -> mean over the local rows.

Here the normalisation is creste_l2norm_rows and everything after the gather is ONE fused kernel pass
(creste_supcon_fwd: online log-sum-exp over the N x Na similarity matrix, never materialised) plus two gradient
passes (creste_supcon_bwd).  The all-gather is the one non-all-reduce collective of the path: NCCL
all_gather forward, reduce-scatter of the column gradients backward."""
import torch
import torch.distributed as dist
from torch import nn

from creste_public_b200 import autograd as ag


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


class _AllGatherRows(torch.autograd.Function):
    """cat(all_gather(x), dim=0) with the gradient of torch.distributed.nn.all_gather: every rank's gradient of
    the gathered tensor is summed and each rank keeps its own slice (a reduce-scatter).  Ranks may hold different
    row counts (the valid-pixel sampling is data dependent): rows are exchanged through a padded buffer."""

    @staticmethod
    def forward(ctx, x, counts):
        world = len(counts)
        nmax = max(counts)
        pad = x.new_zeros(nmax, x.shape[1])
        pad[: x.shape[0]] = x
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        ctx.counts = counts
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)

    @staticmethod
    def backward(ctx, g):
        counts = ctx.counts
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        off = sum(counts[: dist.get_rank()])
        return g[off: off + counts[dist.get_rank()]].clone(), None


def gather_rows(feats, labels):
    """-> (all_feats [Na,D] with gradient, all_labels [Na], self_off = first gathered row of this rank)."""
    if not is_dist_avail_and_initialized() or dist.get_world_size() == 1:
        return feats, labels, 0
    world = dist.get_world_size()
    n = torch.tensor([feats.shape[0]], device=feats.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    counts = [int(v) for v in ns]
    all_feats = _AllGatherRows.apply(feats, counts)
    with torch.no_grad():
        nmax = max(counts)
        lp = labels.new_zeros(nmax)
        lp[: labels.shape[0]] = labels
        lb = [torch.empty_like(lp) for _ in range(world)]
        dist.all_gather(lb, lp)
        all_labels = torch.cat([b[:c] for b, c in zip(lb, counts)], dim=0)
    return all_feats, all_labels, sum(counts[: dist.get_rank()])


class MultiPosConLoss(nn.Module):
    def __init__(self, temperature=0.1, class_weights=None):
        super().__init__()
        self.temperature = temperature
        self.class_weights = class_weights
        # the reference rebuilds its positives mask only when the local row count changes (supcon_loss.py:87-99):
        # two consecutive calls with the same N reuse the FIRST call's labels.  Kept, so that a training run sees
        # the same positives as the reference's.
        self.last_local_batch_size = None
        self._mask_labels = None

    def set_temperature(self, temp=0.1):
        self.temperature = temp

    def forward(self, outputs):
        feats, labels = outputs["feats"], outputs["labels"]           # [N,D], [N]
        cw = None if self.class_weights is None else self.class_weights.to(feats.device)
        D = feats.shape[1]
        if D % 4:
            feats = torch.cat([feats, feats.new_zeros(feats.shape[0], (-D) % 4)], dim=1)
        f = ag.L2NormRowsFn.apply(feats.contiguous().float())
        all_f, all_l, off = gather_rows(f, labels)
        if f.shape[0] != self.last_local_batch_size or self._mask_labels[1].shape[0] != all_l.shape[0]:
            self.last_local_batch_size = f.shape[0]
            self._mask_labels = (labels, all_l)
        ml, mal = self._mask_labels
        loss = ag.SupConFn.apply(f, all_f, ml, mal, off, float(self.temperature), cw)
        return {"loss": loss, "image_loss": loss}


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "models/losses/supcon_loss.py")
