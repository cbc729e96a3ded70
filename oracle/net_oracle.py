"""Functional CPU (torch fp32) restatement of the reference's perception->costmap forward.

TEST INFRASTRUCTURE ONLY (checker).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product package never does.

Parity status: PINNED against the reference's own modules executed in the build container
(tests/test_oracle_cpu.py) and against tests/golden/*.npz (made by oracle/gen_golden.py from
the unmodified reference).  One boundary is UNPINNED BY THE REFERENCE ITSELF: the EfficientNet-B0
trunk lives in the third-party `efficientnet_pytorch` package, which the reference neither
vendors nor pins; `effnet_trunk_endpoints` restates that library's published
algorithm (v0.7.1) and is checked against oracle/ref_shims/efficientnet_shim.py only.

Everything is eval-mode (BatchNorm uses running statistics), written as pure functions over a
flat state dict `sd` that uses the reference's parameter names.  Convolutions are torch CPU
fp32 -- the same arithmetic the reference runs on CPU.

Reference file:line per function:
  effnet_trunk_endpoints   efficientnet_pytorch (external) via creste/models/blocks/effnet.py:83
  effnet_decoder           creste/models/blocks/effnet.py:8-28 (Up), :82-98 (forward)
  depth_completion         creste/models/depth.py:102-158, creste/utils/depth_utils.py:300-313
  dino_head                creste/models/distillation.py:179, creste/models/blocks/conv.py:5-32
  cam2map                  creste/models/blocks/splat_projection.py:131-173, :191-260, :262-354
  bev_decoder              creste/models/blocks/inpainting.py:52-68, :96-109
  reward_fcn / vin_forward creste/models/blocks/conv.py:148-161, creste/models/blocks/vin.py:94-133
  forward                  creste/models/lfd.py:314-330, creste/models/terrainnet.py:272-350
  maxent_irl_loss_value    creste/utils/loss_utils.py:1118-1259 (forward value of the loss terms)
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import c_oracle

B0_STAGES = [  # (repeats, kernel, stride, expand, in, out)
    (1, 3, 1, 1, 32, 16), (2, 3, 2, 6, 16, 24), (2, 5, 2, 6, 24, 40), (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112), (4, 5, 2, 6, 112, 192), (1, 3, 1, 6, 192, 320),
]
PFX_ENC = "backbone.depthcomp.depthcomp.vision_backbone.model."
PFX_DEPTH = "backbone.depthcomp.depthcomp.depth_head.model."
PFX_DINO = "backbone.depthcomp.dino_head.model."
PFX_C2M = "backbone.cam2map."
PFX_BEV = "backbone.bevclassifier."
PFX_VIN = "traversability_head."


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], False, 0.0, eps)


def _swish(x):
    return x * torch.sigmoid(x)


def _same_pad(k, s):
    """Static TF-'SAME' padding as efficientnet_pytorch fixes it at construction (even nominal
    sizes): total = k - s for stride 2 on even sizes, k - 1 for stride 1; split low = total//2."""
    total = (k - 1) if s == 1 else max(k - s, 0)
    lo = total // 2
    return lo, total - lo


def effnet_trunk_endpoints(sd, x, p=PFX_ENC + "trunk."):
    lo, hi = _same_pad(3, 2)
    x = F.conv2d(F.pad(x, (lo, hi, lo, hi)), sd[p + "_conv_stem.weight"], stride=2)
    x = _swish(_bn(x, sd, p + "_bn0", 1e-3))
    endpoints = {}
    prev = x
    idx = 0
    nblocks = sum(s[0] for s in B0_STAGES)
    for (rep, k, s, e, cin, cout) in B0_STAGES:
        for r in range(rep):
            bp = f"{p}_blocks.{idx}."
            stride = s if r == 0 else 1
            inp = x
            if e != 1:
                x = _swish(_bn(F.conv2d(x, sd[bp + "_expand_conv.weight"]), sd, bp + "_bn0", 1e-3))
            lo, hi = _same_pad(k, stride)
            w = sd[bp + "_depthwise_conv.weight"]
            x = F.conv2d(F.pad(x, (lo, hi, lo, hi)), w, stride=stride, groups=w.shape[0])
            x = _swish(_bn(x, sd, bp + "_bn1", 1e-3))
            sq = F.adaptive_avg_pool2d(x, 1)
            sq = _swish(F.conv2d(sq, sd[bp + "_se_reduce.weight"], sd[bp + "_se_reduce.bias"]))
            sq = F.conv2d(sq, sd[bp + "_se_expand.weight"], sd[bp + "_se_expand.bias"])
            x = torch.sigmoid(sq) * x
            x = _bn(F.conv2d(x, sd[bp + "_project_conv.weight"]), sd, bp + "_bn2", 1e-3)
            if stride == 1 and inp.shape[1] == x.shape[1]:
                x = x + inp
            if prev.size(2) > x.size(2):
                endpoints[f"reduction_{len(endpoints) + 1}"] = prev
            elif idx == nblocks - 1:
                endpoints[f"reduction_{len(endpoints) + 1}"] = x
            prev = x
            idx += 1
    return endpoints  # reduction_6 (the 1280-ch head) is never read by EffNet.forward


def _up(sd, p, x1, x2, scale):
    x1 = F.interpolate(x1, scale_factor=scale, mode="bilinear", align_corners=False)
    x = torch.cat([x2, x1], dim=1)
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv.0.weight"], padding=1), sd, p + "conv.1"))
    x = F.relu(_bn(F.conv2d(x, sd[p + "conv.3.weight"], padding=1), sd, p + "conv.4"))
    return x


def up_scales(image_size, downsample=4):
    """effnet.py:52-72: scale factor of each Up stage."""
    scaled = [tuple(image_size)]
    for _ in range(5):
        scaled.insert(0, (scaled[0][0] // 2, scaled[0][1] // 2))
    out, scale, i = [], 32 // downsample, 0
    while scale > 1:
        if not (scaled[i + 1][0] % 2 or scaled[i + 1][1] % 2):
            out.append(2)
        else:
            out.append((scaled[i + 1][0] / scaled[i][0], scaled[i + 1][1] / scaled[i][1]))
        scale //= 2
        i += 1
    return out


def effnet_decoder(sd, x, image_size, p=PFX_ENC):
    ep = effnet_trunk_endpoints(sd, x, p + "trunk.")
    y = ep["reduction_5"]
    for i, sc in enumerate(up_scales(image_size), start=1):
        y = _up(sd, f"{p}up{i}.", y, ep[f"reduction_{5 - i}"], sc)
    return F.conv2d(y, sd[p + "conv.weight"], sd[p + "conv.bias"])


def depth_completion(sd, x, image_size):
    feats = effnet_decoder(sd, x, image_size)
    logits = F.relu(_bn(F.conv2d(feats, sd[PFX_DEPTH + "0.weight"], sd[PFX_DEPTH + "0.bias"],
                                 padding=1), sd, PFX_DEPTH + "1"))
    probs = F.softmax(logits, dim=1)
    vals = torch.linspace(300, 25600, 128).to(logits.dtype).view(1, -1, 1, 1)
    metric = torch.sum(probs * vals, dim=1) / 1000
    return {"depth_preds_logits": logits, "depth_preds_metric": metric,
            "depth_preds_bins": logits.argmax(dim=1), "depth_preds_feats": feats}


def dino_head(sd, feats):
    x = feats
    for i in (0, 3, 6):
        x = F.relu(_bn(F.conv2d(x, sd[f"{PFX_DINO}{i}.weight"], sd[f"{PFX_DINO}{i}.bias"]), sd,
                       f"{PFX_DINO}{i + 1}"))
    return x


def fused_point_features(sd, feats, z):
    """splat_projection.py:152-165: z-MLP, concat, 1x1 fusion conv + BN + ReLU.
    feats [N,256,Hs,Ws], z [N,Hs,Ws] -> [N,96,Hs,Ws]."""
    N, _, Hs, Ws = feats.shape
    zf = z.reshape(-1, 1)
    zf = F.relu(F.linear(zf, sd[PFX_C2M + "z_proj.0.weight"], sd[PFX_C2M + "z_proj.0.bias"]))
    zf = F.relu(F.linear(zf, sd[PFX_C2M + "z_proj.2.weight"], sd[PFX_C2M + "z_proj.2.bias"]))
    zf = zf.view(N, Hs, Ws, -1).permute(0, 3, 1, 2)
    x = torch.cat([feats, zf], dim=1)
    x = F.conv2d(x, sd[PFX_C2M + "vision_fusion.convs.0.weight"],
                 sd[PFX_C2M + "vision_fusion.convs.0.bias"])
    return F.relu(_bn(x, sd, PFX_C2M + "vision_fusion.convs.1"))


def cam2map(sd, depth, feats, p2p, grid=(256, 256)):
    """depth [N,Hs,Ws] (m), feats [N,256,Hs,Ws], p2p [N,4,4] -> bev dict (num_cams=1)."""
    N, Hs, Ws = depth.shape
    rng = sd[PFX_C2M + "point_cloud_range"].numpy()
    vox = sd[PFX_C2M + "voxel_size"].numpy()
    xyz, xy, mask = c_oracle.frustum_to_bev(depth.numpy(), p2p.numpy(), rng, vox)
    z = torch.from_numpy(xyz[:, 2]).view(N, Hs, Ws)
    f = fused_point_features(sd, feats, z)
    f = f * torch.from_numpy(mask).view(N, 1, Hs, Ws).float()
    vol, dens, idx, _ = c_oracle.splat_soft(xy, f.reshape(N, f.shape[1], Hs * Ws).numpy(),
                                            grid[0], grid[1])
    return {"bev_features": torch.from_numpy(vol).view(N, -1, grid[0], grid[1]),
            "bev_densities": torch.from_numpy(dens).view(N, 1, grid[0], grid[1]),
            "bev_coords": torch.from_numpy(xy), "_splat_idx": torch.from_numpy(idx),
            "_fused_feats": f}


def _basic_block(sd, p, x, stride):
    idt = x
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], stride=stride, padding=1), sd, p + "bn1"))
    out = _bn(F.conv2d(out, sd[p + "conv2.weight"], padding=1), sd, p + "bn2")
    if (p + "downsample.0.weight") in sd:
        idt = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1")
    return F.relu(out + idt)


def bev_decoder(sd, bev, prefixes=("inpainting_sam", "inpainting_sam_dynamic", "elevation")):
    p = PFX_BEV
    x = F.relu(_bn(F.conv2d(bev, sd[p + "conv1.weight"], stride=2, padding=3), sd, p + "bn1"))
    x1 = _basic_block(sd, p + "layer1.1.", _basic_block(sd, p + "layer1.0.", x, 1), 1)
    x = _basic_block(sd, p + "layer2.1.", _basic_block(sd, p + "layer2.0.", x1, 2), 1)
    x = _basic_block(sd, p + "layer3.1.", _basic_block(sd, p + "layer3.0.", x, 2), 1)
    out = {}
    for h, name in enumerate(prefixes):
        hp = f"{p}out_heads.{h}."
        y = _up(sd, hp + "up1.", x, x1, 4)
        y = F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=False)
        y = F.relu(_bn(F.conv2d(y, sd[hp + "up2.1.weight"], padding=1), sd, hp + "up2.2"))
        out[f"{name}_preds"] = F.conv2d(y, sd[hp + "proj.weight"], sd[hp + "proj.bias"])
        out[f"{name}_features"] = y
    return out


def _conv_layer(sd, p, x, bn=True, relu=True):
    w = sd[p + "conv.weight"]
    x = F.conv2d(x, w, padding=w.shape[-1] // 2)
    if bn:
        x = _bn(x, sd, p + "norm")
    return F.relu(x) if relu else x


def reward_fcn(sd, x, p=PFX_VIN + "r."):
    x = _conv_layer(sd, p + "prepool.1.", _conv_layer(sd, p + "prepool.0.", x))
    skip = _conv_layer(sd, p + "skip.1.", _conv_layer(sd, p + "skip.0.", x))
    t = F.max_pool2d(x, 2, 2)
    t = F.relu(_bn(_conv_layer(sd, p + "trunk.1.", t, bn=False), sd, p + "trunk.2"))
    t = F.relu(_bn(_conv_layer(sd, p + "trunk.4.", t, bn=False), sd, p + "trunk.5"))
    t = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    return _conv_layer(sd, p + "postpool.0.", torch.cat([t, skip], dim=1))


def vin_forward(sd, feat_map, ds=2,
                keys=("inpainting_sam_preds", "inpainting_sam_dynamic_preds", "elevation_preds")):
    iv = torch.cat([feat_map[k] for k in keys], dim=1)
    Ho, Wo = iv.shape[-2:]
    iv = F.max_pool2d(iv, ds, ds)
    iv = iv[:, :, : iv.shape[2] // 2, :]
    r = reward_fcn(sd, iv)
    full = torch.zeros(iv.shape[0], 1, Ho, Wo)
    full[:, :, : Ho // 2, :] = F.interpolate(r, size=(Ho // 2, Wo), mode="bilinear",
                                             align_corners=False)
    return {"traversability_preds": r, "traversability_preds_full": full, "input_view": iv}


@torch.no_grad()
def forward(sd, rgbd, p2p, encoder_fp64=False):
    """rgbd [B,1,4,H,W], p2p [B,1,4,4] -> the reference's output dict (solve_mdp=False).

    encoder_fp64=True evaluates the RGB-D encoder + depth head in float64 (i.e. exactly) and
    hands the rounded result to the unchanged fp32 rest: the distance between that output and
    the plain fp32 one measures how far the reference's OWN rounding noise moves each tensor
    (the conditioning yardstick of the end-to-end parity test)."""
    B, V, C, H, W = rgbd.shape
    assert V == 1
    x = rgbd.view(B, C, H, W)
    if encoder_fp64:
        sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()
                if k.startswith("backbone.depthcomp.")}
        out = depth_completion(sd64, x.double(), (H, W))
        out = {k: (v.float() if v.is_floating_point() else v) for k, v in out.items()}
    else:
        out = depth_completion(sd, x, (H, W))
    out["dino_pe_feats"] = dino_head(sd, out["depth_preds_feats"]).unsqueeze(1)
    bev = cam2map(sd, out["depth_preds_metric"], out["depth_preds_feats"], p2p.view(B, 4, 4))
    out.update({k: v for k, v in bev.items()})
    out.update(bev_decoder(sd, out["bev_features"]))
    out.update(vin_forward(sd, out))
    return out


from synth_data import trapezoid_fov_mask  # noqa: E402,F401  (shared synthetic-input generator)


def maxent_irl_loss_value(exp_svf, expert_rc, fov_mask_full, reward, cf_list, map_ds=2,
                          map_sz=(64, 128), alpha=0.5, use_fov_mask=True):
    """Forward value of the visitation term of MaxEntIRLLoss (loss_utils.py:1118-1203).
    exp_svf [B,H,W]; expert_rc [B,T,2]; fov_mask_full [B,Ho,Wo] bool; reward [B,H,W];
    cf_list: per-sample None or dict(trajectories f64 [N,T,2], rank [N]).  Returns
    (visitation_loss, mean_exp_svf_rewards, mean_svf_rewards) as python floats."""
    exp_svf = torch.as_tensor(exp_svf).clone()
    reward = torch.as_tensor(reward)
    B, H, W = exp_svf.shape
    fm = torch.as_tensor(fov_mask_full)
    Ho, Wo = fm.shape[-2:]
    fm = F.interpolate(fm.unsqueeze(1).byte(), size=(Ho // 2, Wo // 2), mode="nearest")
    fm = fm[:, 0, 0:H, 0:W].bool()
    svf = torch.from_numpy(c_oracle.expert_visitation(np.asarray(expert_rc), map_ds, map_sz[0],
                                                      map_sz[1], False))
    if use_fov_mask:
        svf = svf * fm.float()
        exp_svf = exp_svf * fm.float()
    svf = svf / (svf.sum(dim=(1, 2), keepdim=True) + 1e-5)
    exp_svf = exp_svf / (exp_svf.sum(dim=(1, 2), keepdim=True) + 1e-5)
    for i, cf in enumerate(cf_list or []):
        if cf is None:
            continue
        bad = np.asarray(cf["trajectories"])[np.asarray(cf["rank"]) > 0]
        if bad.shape[0] == 0:
            continue
        c = torch.from_numpy(c_oracle.expert_visitation(bad, map_ds, map_sz[0], map_sz[1], True))
        c = c.sum(dim=0)
        c = c / (c.sum() + 1e-5)
        exp_svf[i] = alpha * c + (1 - alpha) * exp_svf[i]
    if use_fov_mask:
        reward = reward * fm.float()
    a = (exp_svf * reward).sum(dim=(1, 2)).mean()
    b = (svf * reward).sum(dim=(1, 2)).mean()
    return float(a - b), float(a), float(b)
