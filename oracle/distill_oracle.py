"""Stage-1 (distillation) training step oracle.  TEST INFRASTRUCTURE -- never imported by the product.

`reference_step` drives the UNMODIFIED reference (creste/models/distillation.py DistillationBackbone
in train mode + creste/utils/loss_utils.py LossManager with the three losses of
configs/model/distillation/effnet_ds2_dinov2_128.yaml:72-88 + torch.optim.Adam as
creste/train_pefree.py:176-181 builds it) under the import shims -- build container only.
`port_step` restates the same step on plain torch CPU modules so that it can run on the GPU box:

  PortDistillation  distillation.py:145-207 / depth.py:102-158 / effnet.py:8-98 / conv.py:5-32 over the
                    EfficientNet-B0 restatement of oracle/ref_shims/efficientnet_shim.py (state-dict
                    names equal the reference's, so state dicts are interchangeable)
  port_losses       loss_utils.py:477-527 (CrossEntropyDepth), :530-573 (SmoothL1Depth on the int64
                    bins: a value without a gradient), :606-647 (MSELoss), depth_utils.py:346-383

Parity status: PINNED -- port and reference agree on loss, gradients and post-Adam parameters
(tests/test_distill_cpu.py::test_port_matches_reference) and tests/golden/distill_step.npz holds the
reference's own outputs for the seeded case.  The EfficientNet-B0 trunk itself is the un-vendored
third-party boundary described in oracle/ref_shims/efficientnet_shim.py (unpinned by the reference).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import synth
from .ref_shims import efficientnet_shim as effs

DISC = {"mode": "UD", "num_bins": 128, "depth_min": 300, "depth_max": 25600}
W_CE, W_SL1, W_MSE, BETA = 0.5, 0.1, 1.0, 0.5


# ------------------------------------------------------------------------------------ the port
class _Up(nn.Module):
    def __init__(self, inC, outC, sf=2):
        super().__init__()
        self.up = nn.Upsample(scale_factor=sf, mode="bilinear", align_corners=False)
        self.conv = nn.Sequential(
            nn.Conv2d(inC, outC, 3, padding=1, bias=False), nn.BatchNorm2d(outC), nn.ReLU(inplace=True),
            nn.Conv2d(outC, outC, 3, padding=1, bias=False), nn.BatchNorm2d(outC), nn.ReLU(inplace=True))

    def forward(self, x1, x2):
        return self.conv(torch.cat([x2, self.up(x1)], dim=1))


class _EffNet(nn.Module):
    """effnet.py:31-98 for efficientnet-b0, inC=4, downsample=4 (three Up stages, x2 each)."""

    def __init__(self, image_size, inC=4, outC=256):
        super().__init__()
        self.trunk = effs.EfficientNet.from_pretrained("efficientnet-b0")
        self.trunk.set_swish(memory_efficient=False)
        self.trunk._conv_stem = effs.get_same_padding_conv2d(tuple(image_size))(inC, 32, kernel_size=3,
                                                                                stride=2, bias=False)
        ch = [320, 112, 40, 24]
        C = ch[0]
        for i in (1, 2, 3):
            C += ch[i]
            setattr(self, f"up{i}", _Up(C, C, 2))
        self.conv = nn.Conv2d(C, outC, 1)

    def forward(self, x):
        ep = self.trunk.extract_endpoints(x)
        y = ep["reduction_5"]
        for i in (1, 2, 3):
            y = getattr(self, f"up{i}")(y, ep[f"reduction_{5 - i}"])
        return self.conv(y)


class _Holder(nn.Module):
    def __init__(self, name, mod):
        super().__init__()
        setattr(self, name, mod)


def _stack(dims, kernels, pads):
    m = []
    for i, k in enumerate(kernels):
        m += [nn.Conv2d(dims[i], dims[i + 1], k, padding=pads[i]), nn.BatchNorm2d(dims[i + 1]), nn.ReLU()]
    return nn.Sequential(*m)


class PortDistillation(nn.Module):
    def __init__(self, image_size):
        super().__init__()
        dc = nn.Module()
        dc.vision_backbone = _Holder("model", _EffNet(image_size))
        dc.depth_head = _Holder("model", _stack([256, 128], [3], [1]))
        self.depthcomp = dc
        self.dino_head = _Holder("model", _stack([256, 128, 128, 128], [1, 1, 1], [0, 0, 0]))

    def forward(self, rgbd):
        B, V, Cc, H, W = rgbd.shape
        x = rgbd.view(B * V, Cc, H, W)
        feats = self.depthcomp.vision_backbone.model(x)
        logits = self.depthcomp.depth_head.model(feats)
        probs = F.softmax(logits, dim=1)
        vals = torch.linspace(DISC["depth_min"], DISC["depth_max"], DISC["num_bins"]).view(1, -1, 1, 1)
        dino = self.dino_head.model(feats)
        return {"depth_preds_logits": logits, "depth_preds_metric": torch.sum(probs * vals, dim=1) / 1000,
                "depth_preds_bins": torch.argmax(logits, dim=1), "depth_preds_feats": feats,
                "dino_pe_feats": dino.view(B, V, *dino.shape[1:])}


def _bin_depths(d):
    bin_size = (DISC["depth_max"] - DISC["depth_min"]) / DISC["num_bins"]
    idx = (d - DISC["depth_min"]) / bin_size
    mask = (idx < 0) | (idx > DISC["num_bins"]) | (~torch.isfinite(idx))
    idx = idx.clone()
    idx[mask] = DISC["num_bins"]
    return idx.type(torch.int64)


def port_losses(outputs, inputs):
    """-> (total, {name: value}) with the names LossManager gives the shipped config."""
    lab = inputs["depth_label"]
    B, S, H, W = lab.shape
    gt = lab.view(B * S, H, W)
    gt_bin = _bin_depths(gt)
    logits = outputs["depth_preds_logits"]
    flat_gt = gt_bin.flatten(1, 2).long()
    valid = flat_gt != DISC["num_bins"]
    flat_pred = logits.permute(0, 2, 3, 1).flatten(1, 2)
    ce = F.cross_entropy(flat_pred[valid, :], flat_gt[valid])
    acc = (flat_gt[valid] == flat_pred[valid, :].argmax(1)).sum() / flat_gt[valid].numel()
    v2 = gt_bin != DISC["num_bins"]
    sl1 = F.smooth_l1_loss(outputs["depth_preds_bins"][v2].float(), (gt / 1000.0)[v2].float(), beta=BETA)
    pred, tgt = outputs["dino_pe_feats"], inputs["fimg_label"]
    Bq, V, Z, Hh, Ww = tgt.shape
    tgt = tgt.permute(0, 1, 3, 4, 2).reshape(Bq * V * Hh * Ww, Z)
    pred = pred.permute(0, 1, 3, 4, 2).reshape(Bq * V * Hh * Ww, Z)
    ok = ~torch.isinf(tgt)
    mse = F.mse_loss(pred[ok], tgt[ok])
    total = W_CE * ce + W_SL1 * sl1 + W_MSE * mse
    return total, {"CrossEntropyDepth/depth/cls_loss": ce, "SmoothL1Depth/depth/reg_loss": sl1,
                   "MSELoss/loss": mse, "CrossEntropyDepth/depth/acc": acc}


# ------------------------------------------------------------------------------------ the case
def make_case(seed=5, B=2, image_size=(64, 96)):
    """Seeded inputs + parameters of one stage-1 step (shrunk)."""
    H, W = image_size
    g = np.random.default_rng(7000 + seed)
    net = PortDistillation(image_size)
    sd = synth.seeded_state_dict(net.state_dict(), seed)
    rgb = g.random((B, 1, 3, H, W)).astype(np.float32)
    depth = g.uniform(300, 25600, (B, 1, 1, H, W)).astype(np.float32)
    depth[g.random(depth.shape) > 0.04] = 0.0                  # ~4 % LiDAR fill, millimetres
    lab = g.uniform(300, 25600, (B, 1, H // 4, W // 4)).astype(np.float32)
    lab[g.random(lab.shape) < 0.2] = 0.0                        # 20 % invalid
    lab[0, 0, 0, 0] = 25600.0                                   # == depth_max -> bin 128 -> invalid
    lab[0, 0, 0, 1] = 300.0                                     # == depth_min -> bin 0 (valid)
    fimg = g.standard_normal((B, 1, 128, H // 4, W // 4)).astype(np.float32)
    fimg[g.random(fimg.shape) < 0.01] = np.inf                  # masked targets (loss_utils.py:643)
    return {"image_size": tuple(image_size), "state_dict": sd, "seed": seed,
            "image": torch.from_numpy(np.concatenate([rgb, depth], axis=2)),
            "depth_label": torch.from_numpy(lab), "fimg_label": torch.from_numpy(fimg)}


def _collect(model, losses, total, outputs):
    out = {"loss": np.float32(total.detach().cpu()),
           "logits": outputs["depth_preds_logits"].detach().cpu().numpy().copy(),
           "dino": outputs["dino_pe_feats"].detach().cpu().numpy().copy()}
    out.update({k: np.float32(v.detach().cpu()) for k, v in losses.items()})
    out["grads"] = {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters()
                    if p.grad is not None}
    return out


def _finish(model, out):
    out["params"] = {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}
    return out


def port_step(case, lr=5e-4):
    """One training step of the port: forward (train mode, drop-connect drawn from torch's CPU generator
    seeded with case['seed']) -> losses -> backward -> Adam."""
    model = PortDistillation(case["image_size"])
    model.load_state_dict(case["state_dict"])
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999))
    torch.manual_seed(case["seed"])
    outputs = model(case["image"].clone())
    total, losses = port_losses(outputs, {"depth_label": case["depth_label"].clone(),
                                          "fimg_label": case["fimg_label"].clone()})
    opt.zero_grad()
    total.backward()
    out = _collect(model, losses, total, outputs)
    opt.step()
    return _finish(model, out)


def port_grads_fp64(case, relu_hook=None):
    """The same step evaluated in float64 (same drop-connect masks): the yardstick that separates an
    implementation's error from the fp32 rounding noise of the reference itself.  -> (grads, loss)
    relu_hook(i, u) -> u' | None is called on the input of the i-th ReLU in execution order (the six of the Up
    blocks, the depth head's, the dino head's three): parity tests read the oracle's ReLU masks through it and break
    near-ties (a pre-activation closer to zero than fp32 rounding noise) the way the implementation under test did."""
    model = PortDistillation(case["image_size"])
    model.load_state_dict(case["state_dict"])
    model.double().train()
    if relu_hook is not None:
        count = [0]

        def pre(mod, args):
            i = count[0]
            count[0] += 1
            return relu_hook(i, args[0])

        for m in model.modules():
            if isinstance(m, nn.ReLU):
                m.register_forward_pre_hook(pre)
    orig = effs.drop_connect

    def drop_connect32(inputs, p, training):          # the fp32 run's uniforms, whatever the dtype
        if not training:
            return inputs
        keep = 1 - p
        r = keep + torch.rand([inputs.shape[0], 1, 1, 1], dtype=torch.float32).to(inputs.dtype)
        return inputs / keep * torch.floor(r)
    effs.drop_connect = drop_connect32
    try:
        torch.manual_seed(case["seed"])
        outputs = model(case["image"].double())
        total, _ = port_losses(outputs, {"depth_label": case["depth_label"].double(),
                                         "fimg_label": case["fimg_label"].double()})
        total.backward()
    finally:
        effs.drop_connect = orig
    return ({k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
            float(total.detach()))


def reference_step(case, lr=5e-4):
    """The unmodified reference modules (build container only)."""
    from . import ref_harness as rh
    mods = rh.ref_modules()
    from omegaconf import OmegaConf
    cfg = rh.compose_cfgs(image_size=case["image_size"])["distill"]
    model = mods["distillation"].DistillationBackbone(OmegaConf.create(cfg))
    model.load_state_dict(case["state_dict"])
    model.train()
    loss_mgr = mods["loss_utils"].LossManager(OmegaConf.create(cfg))
    opt = torch.optim.Adam(model.parameters(), lr=lr, betas=(0.9, 0.999))
    torch.manual_seed(case["seed"])
    outputs = model(case["image"].clone())
    inputs = {"image": case["image"], "depth_label": case["depth_label"].clone(),
              "fimg_label": case["fimg_label"].clone()}
    merged = mods["train_utils"].merge_dict(("inputs", inputs), ("outputs", outputs))
    loss_dict, meta = loss_mgr(merged)
    total = sum(w * v for w, v in loss_dict.values())
    losses = {k: v for k, (w, v) in loss_dict.items()}
    losses.update(meta)
    opt.zero_grad()
    total.backward()
    out = _collect(model, losses, total, outputs)
    opt.step()
    return _finish(model, out)
