"""Train-mode BEV decoder step -- oracle groundwork for the stage-2 ("next") training
surface.  TEST INFRASTRUCTURE; nothing in the product uses it yet.

`reference_step` drives the unmodified reference `InpaintingResNet18MultiHead`
(creste/models/blocks/inpainting.py:70-109: 7x7 stride-2 stem, torchvision ResNet-18 layer1..3 with their
stride-2 BasicBlocks, three DeconvHeads :52-68) in train mode under the shims (build container only);
`port_step` restates it on torchvision + plain torch modules so that it runs on the GPU box.  The scalar that is
differentiated is  sum_h <preds_h, P_h> + 1e-2 <features_h, F_h>  with seeded P, F, so that every parameter of
the decoder (both strided convolutions of layer2 / layer3 and their 1x1 downsample branches included) receives a
gradient.  Port and reference agree bit for bit (tests/test_oracle_cpu.py::test_bev_port_matches_reference);
tests/golden/bev_step.npz holds the reference's loss and the L2 norm of every gradient."""
import numpy as np
import torch
import torchvision
from torch import nn

from . import synth

NUM_CLASSES = (32, 6, 2)
PREFIXES = ("inpainting_sam", "inpainting_sam_dynamic", "elevation")


class _Up(nn.Module):
    def __init__(self, inC, outC, scale_factor):
        super().__init__()
        self.up = nn.Upsample(scale_factor=scale_factor, mode="bilinear", align_corners=False)
        self.conv = nn.Sequential(
            nn.Conv2d(inC, outC, 3, padding=1, bias=False), nn.BatchNorm2d(outC), nn.ReLU(inplace=True),
            nn.Conv2d(outC, outC, 3, padding=1, bias=False), nn.BatchNorm2d(outC), nn.ReLU(inplace=True))

    def forward(self, x1, x2):
        return self.conv(torch.cat([x2, self.up(x1)], dim=1))


class _Head(nn.Module):
    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.up1 = _Up(in_ch, 256, 4)
        self.up2 = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
                                 nn.Conv2d(256, 128, 3, padding=1, bias=False), nn.BatchNorm2d(128),
                                 nn.ReLU(inplace=True))
        self.proj = nn.Conv2d(128, out_ch, 1)

    def forward(self, x1, x2):
        x = self.up2(self.up1(x1, x2))
        return self.proj(x), x


class PortBEVDecoder(nn.Module):
    def __init__(self, num_input_features=96, num_classes=NUM_CLASSES):
        super().__init__()
        trunk = torchvision.models.resnet.resnet18(weights=None, zero_init_residual=True)
        self.conv1 = nn.Conv2d(num_input_features, 64, 7, stride=2, padding=3, bias=False)
        self.bn1, self.relu = trunk.bn1, trunk.relu
        self.layer1, self.layer2, self.layer3 = trunk.layer1, trunk.layer2, trunk.layer3
        self.out_heads = nn.ModuleList([_Head(64 + 256, n) for n in num_classes])

    def forward(self, x):
        x1 = self.layer1(self.relu(self.bn1(self.conv1(x))))
        x = self.layer3(self.layer2(x1))
        return [head(x, x1) for head in self.out_heads]


def make_case(seed=9, B=2, H=32, W=32, Cin=96):
    g = np.random.default_rng(8000 + seed)
    sd = synth.seeded_state_dict(PortBEVDecoder(Cin).state_dict(), seed)
    bev = (g.standard_normal((B, Cin, H, W)) * (g.random((B, 1, H, W)) < 0.3)).astype(np.float32)   # sparse, like a splat
    P = [g.standard_normal((B, n, H, W)).astype(np.float32) for n in NUM_CLASSES]
    F = [g.standard_normal((B, 128, H, W)).astype(np.float32) for _ in NUM_CLASSES]
    return {"state_dict": sd, "bev": torch.from_numpy(bev), "P": [torch.from_numpy(p) for p in P],
            "F": [torch.from_numpy(f) for f in F]}


def _finish(model, outs, case):
    loss = sum((p * P).sum() + 1e-2 * (f * Fw).sum() for (p, f), P, Fw in zip(outs, case["P"], case["F"]))
    loss.backward()
    return {"loss": np.float32(loss.detach()),
            "preds0": outs[0][0].detach().numpy().copy(),
            "grads": {k: p.grad.numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
            "buffers": {k: v.numpy().copy() for k, v in model.state_dict().items() if "running" in k}}


def port_step(case):
    model = PortBEVDecoder(case["bev"].shape[1])
    model.load_state_dict(case["state_dict"])
    model.train()
    return _finish(model, model(case["bev"].clone()), case)


def reference_step(case):
    from . import ref_harness as rh
    rh.ref_modules()
    import creste.models.blocks.inpainting as inp
    model = inp.InpaintingResNet18MultiHead(case["bev"].shape[1], list(NUM_CLASSES), norm_layer="batch_norm",
                                            input_key="bev_features", output_prefix=list(PREFIXES))
    model.load_state_dict(case["state_dict"])
    model.train()
    out = model({"bev_features": case["bev"].clone()})
    outs = [(out[f"{p}_preds"], out[f"{p}_features"]) for p in PREFIXES]
    return _finish(model, outs, case)
