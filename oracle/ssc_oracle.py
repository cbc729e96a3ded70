"""Stage-2 (train_ssc.py) training-graph oracle.  TEST INFRASTRUCTURE -- never imported by the product.

`reference_step` drives the UNMODIFIED reference `TerrainNet` (creste/models/terrainnet.py:272-350: RGB-D backbone ->
depth-guided frustum->BEV splat -> ResNet-18 BEV decoder with three DeconvHeads) in TRAIN mode under the import
shims -- build container only.  `port_step` restates the same graph on plain torch modules so that it can run on the
GPU box:

  PortDistillation   oracle/distill_oracle.py (pinned to the reference bit for bit)
  PortCam2Map        creste/models/blocks/splat_projection.py:19-51 (un-projection), :131-173 (z-MLP, fusion conv +
                     BatchNorm + ReLU, bounds mask), :175-189 (voxel coordinates), :262-354 (4-tap scatter_add_ splat,
                     mean normalisation) -- all differentiable torch ops, so autograd reaches the image features, the
                     z-MLP / fusion parameters and, through the tap weights and z, the predicted depth
  PortBEVDecoder     oracle/bev_oracle.py (pinned to the reference bit for bit)

State-dict names equal the reference's (tests/test_stage2_cpu.py loads one state dict into all three).

The differentiated scalar is  sum_k <outputs[k], W_k> * s_k  over every floating-point output of the forward (depth
logits, soft-argmax depth, dino features, BEV features / densities, the three heads' predictions and features) with
seeded cotangents W_k, so that every parameter of the model and every backward path (splat -> features, splat ->
voxel coordinates -> depth -> logits, strided convs, up-sampling adjoints) receives a gradient.  The stage-2 LOSSES
(SupPixelConLoss, CrossEntropy, SmoothL1, ...) are checked separately (tests/test_stage2_losses_*.py)."""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import bev_oracle, distill_oracle, eager_oracle
from .distill_oracle import _Holder

PREFIXES = bev_oracle.PREFIXES
OUT_KEYS = (("depth_preds_logits", 1e-2), ("depth_preds_metric", 1e-1), ("dino_pe_feats", 1e-2), ("bev_features", 1e-1),
            ("bev_densities", 1e-1)) + tuple((f"{p}_preds", 1e-2) for p in PREFIXES) + \
    tuple((f"{p}_features", 1e-3) for p in PREFIXES)


class PortCam2Map(nn.Module):
    def __init__(self, pc_range=(-12.8, -12.8, -2, 12.8, 12.8, 1), voxel=(0.1, 0.1, 3)):
        super().__init__()
        self.register_buffer("point_cloud_range", torch.tensor(pc_range, dtype=torch.float32))
        self.register_buffer("max_bound", self.point_cloud_range[3:].reshape(1, -1))
        self.register_buffer("min_bound", self.point_cloud_range[:3].reshape(1, -1))
        self.register_buffer("voxel_size", torch.tensor(voxel, dtype=torch.float32))
        self.register_buffer("grid_size", ((self.point_cloud_range[3:] - self.point_cloud_range[:3]) / self.voxel_size).long())
        self.register_buffer("lidar2map", torch.tensor([
            [0, -1, 0, -self.min_bound[0, 0]], [-1, 0, 0, -self.min_bound[0, 1]],
            [0, 0, -1, -self.min_bound[0, 2]], [0, 0, 0, 1]]).float())
        self.z_proj = nn.Sequential(nn.Linear(1, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU())
        self.vision_fusion = _Holder("convs", nn.Sequential(nn.Conv2d(288, 96, 1), nn.BatchNorm2d(96), nn.ReLU()))

    def forward(self, depth, feats, p2p):
        M, Hs, Ws = depth.shape
        xyz = eager_oracle.camera_to_world(depth, p2p)                                   # [M,3,Hs,Ws]
        zf = self.z_proj(xyz[:, 2].reshape(-1, 1)).view(M, Hs, Ws, -1).permute(0, 3, 1, 2)
        f = self.vision_fusion.convs(torch.cat([feats, zf], dim=1))
        pts = xyz.permute(0, 2, 3, 1).reshape(M, Hs * Ws, 3)
        mask = torch.all((pts < self.max_bound) & (pts >= self.min_bound), dim=2, keepdim=True)
        f = f * mask.view(M, Hs, Ws, 1).permute(0, 3, 1, 2)
        hom = torch.cat([pts, torch.ones_like(pts[:, :, :1])], dim=2)
        hom = (self.lidar2map @ hom.permute(0, 2, 1)).permute(0, 2, 1)
        xy = hom[:, :, :2] / self.voxel_size[:2]
        H, W = int(self.grid_size[0]), int(self.grid_size[1])
        vol, dens = eager_oracle.splat_soft(xy, f.reshape(M, f.shape[1], Hs * Ws), H, W)
        return {"bev_features": vol.view(M, -1, H, W), "bev_densities": dens.view(M, H, W, 1).permute(0, 3, 1, 2),
                "bev_coords": xy}


class PortTerrainNet(nn.Module):
    def __init__(self, image_size):
        super().__init__()
        self.depthcomp = distill_oracle.PortDistillation(image_size)
        self.cam2map = PortCam2Map()
        self.bevclassifier = bev_oracle.PortBEVDecoder(96)

    def forward(self, x):
        rgbd, p2p = x[:2]
        B = rgbd.shape[0]
        out = dict(self.depthcomp(rgbd))
        out.update(self.cam2map(out["depth_preds_metric"], out["depth_preds_feats"], p2p.view(B, 4, 4)))
        for (pred, feat), p in zip(self.bevclassifier(out["bev_features"]), PREFIXES):
            out[f"{p}_preds"], out[f"{p}_features"] = pred, feat
        return out


def make_case(template_sd, seed=11, B=2, image_size=(64, 96)):
    """Seeded parameters (names / shapes from `template_sd`, a TerrainNet state dict), inputs and cotangents."""
    from . import synth
    H, W = image_size
    g = np.random.default_rng(9100 + seed)
    sd = synth.seeded_state_dict(template_sd, seed, "soft")
    rgbd, p2p = synth.net_inputs(H, W, B, seed=seed)
    cot = {}
    shapes = {"depth_preds_logits": (B, 128, H // 4, W // 4), "depth_preds_metric": (B, H // 4, W // 4),
              "dino_pe_feats": (B, 1, 128, H // 4, W // 4), "bev_features": (B, 96, 256, 256),
              "bev_densities": (B, 1, 256, 256)}
    for p, n in zip(PREFIXES, bev_oracle.NUM_CLASSES):
        shapes[f"{p}_preds"] = (B, n, 256, 256)
        shapes[f"{p}_features"] = (B, 128, 256, 256)
    for k, _ in OUT_KEYS:
        cot[k] = torch.from_numpy(g.standard_normal(shapes[k]).astype(np.float32))
    return {"state_dict": sd, "image_size": tuple(image_size), "seed": seed, "image": rgbd, "p2p": p2p, "cot": cot}


def scalar(outputs, case, dtype=None):
    tot = 0.0
    for k, s in OUT_KEYS:
        w = case["cot"][k].to(outputs[k].device)
        tot = tot + s * (outputs[k] * (w if dtype is None else w.to(dtype))).sum()
    return tot


def _finish(model, outputs, total):
    total.backward()
    return {"loss": np.float64(total.detach().double().cpu()),
            "grads": {k: p.grad.detach().cpu().numpy().copy() for k, p in model.named_parameters() if p.grad is not None},
            "buffers": {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items() if "running" in k},
            "outputs": {k: outputs[k].detach().cpu().numpy().copy() for k in ("depth_preds_metric", "bev_densities")},
            "samples": {k: outputs[k].detach().cpu().numpy()[..., ::4, ::4].copy()
                        for k in ("bev_features", "inpainting_sam_preds", "elevation_preds")}}


def _seed_drop_connect(case):
    torch.manual_seed(case["seed"])            # the EfficientNet drop-connect uniforms (CPU generator)


def port_step(case, dtype=torch.float32):
    model = PortTerrainNet(case["image_size"])
    model.load_state_dict(case["state_dict"])
    model = model.to(dtype).train()
    _seed_drop_connect(case)
    if dtype == torch.float64:
        from .ref_shims import efficientnet_shim as effs
        orig = effs.drop_connect

        def drop_connect32(inputs, p, training):          # the fp32 run's uniforms, whatever the dtype
            if not training:
                return inputs
            keep = 1 - p
            r = keep + torch.rand([inputs.shape[0], 1, 1, 1], dtype=torch.float32).to(inputs.dtype)
            return inputs / keep * torch.floor(r)
        effs.drop_connect = drop_connect32
        try:
            out = model((case["image"].to(dtype), case["p2p"].to(dtype)))
        finally:
            effs.drop_connect = orig
    else:
        out = model((case["image"].clone(), case["p2p"].clone()))
    return _finish(model, out, scalar(out, case, dtype))


def reference_step(case):
    from . import ref_harness as rh
    mods = rh.ref_modules()
    from omegaconf import OmegaConf
    cfg = rh.compose_cfgs(image_size=case["image_size"])["ssc"]
    model = mods["terrainnet"].TerrainNet(OmegaConf.create(cfg))
    model.load_state_dict(case["state_dict"])
    model.train()
    _seed_drop_connect(case)
    out = model((case["image"].clone(), case["p2p"].clone(), None))
    return _finish(model, out, scalar(out, case))
