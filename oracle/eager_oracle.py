"""Device-agnostic PyTorch-eager restatement of the reference's perception->costmap forward: the BASELINE the
north_star's ">= 20x the reference single-GPU PyTorch frames/sec" target is measured against.

TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's `gpu_eager_baseline` leg runs it on cuda:0 (the reference itself is
Python and cannot travel to the GPU box); tests compare it with oracle/net_oracle.py on the CPU.  The product never
imports it.

It issues the same library calls as the reference does on a GPU -- cuDNN convolutions / BatchNorm, `F.interpolate`,
the `torch.meshgrid` + `bmm` un-projection (creste/models/blocks/splat_projection.py:19-51), the 4-tap
`scatter_add_` splat with a random out-of-bounds target (:262-354), `max_pool2d` + the reward FCN
(creste/models/blocks/vin.py:94-133) -- over a flat state dict with the reference's parameter names.  The conv
stacks reuse the functional pieces of net_oracle (pure torch, device-agnostic)."""
import torch
import torch.nn.functional as F

from . import net_oracle as no


def depth_completion(sd, x, image_size):
    feats = no.effnet_decoder(sd, x, image_size)
    logits = F.relu(no._bn(F.conv2d(feats, sd[no.PFX_DEPTH + "0.weight"], sd[no.PFX_DEPTH + "0.bias"], padding=1),
                           sd, no.PFX_DEPTH + "1"))
    probs = F.softmax(logits, dim=1)
    vals = torch.linspace(300, 25600, 128, device=logits.device).view(1, -1, 1, 1)
    metric = torch.sum(probs * vals, dim=1) / 1000
    return {"depth_preds_logits": logits, "depth_preds_metric": metric, "depth_preds_bins": logits.argmax(dim=1),
            "depth_preds_feats": feats}


def camera_to_world(depth, p2p):
    """splat_projection.py:19-51 (the pixel grid is rebuilt and uploaded on every call, as in the reference)."""
    BN, H, W = depth.shape
    u, v = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    cam = torch.tile(torch.stack([u, v, torch.ones_like(u)], dim=0), (BN, 1, 1, 1)).to(depth.device)
    cam = cam * depth.view(BN, 1, H, W)
    cam = torch.cat([cam, torch.ones_like(depth.view(BN, 1, H, W))], dim=1)
    return torch.bmm(p2p, cam.flatten(start_dim=2)).view(BN, 4, H, W)[:, :3]


def splat_soft(xy, feats, H, W, min_weight=1.0):
    """splat_projection.py:262-354, scatter_mode='mean': xy [B,P,2], feats [B,F,P]."""
    ba, fd, n_points = feats.shape
    n_voxels = H * W
    XY = xy.floor().long()
    rXY = xy - XY.type_as(xy)
    X, Y = XY.split(1, dim=2)
    rX, rY = rXY.split(1, dim=2)
    rand_idx = X.new_zeros(X.shape).random_(0, n_voxels)
    dens = feats.new_zeros(ba, n_voxels, 1)
    vol = feats.new_zeros(ba, fd, n_voxels)
    for xdiff in (0, 1):
        X_ = X + xdiff
        wX = (1 - xdiff) + (2 * xdiff - 1) * rX
        for ydiff in (0, 1):
            Y_ = Y + ydiff
            wY = (1 - ydiff) + (2 * ydiff - 1) * rY
            w = wX * wY
            valid = ((0 <= X_) * (X_ < W) * (0 <= Y_) * (Y_ < H)).long()
            idx = Y_ * W + X_
            idx_valid = idx * valid + rand_idx * (1 - valid)
            w_valid = w * valid.type_as(w)
            dens.scatter_add_(1, idx_valid, w_valid)
            vol.scatter_add_(2, idx_valid.view(ba, 1, n_points).expand_as(feats), w_valid.view(ba, 1, n_points) * feats)
    vol = vol / dens.view(ba, 1, n_voxels).clamp(min_weight)
    return vol, dens


def cam2map(sd, depth, feats, p2p, grid=(256, 256)):
    N, Hs, Ws = depth.shape
    xyz = camera_to_world(depth, p2p)                                          # [N,3,Hs,Ws]
    f = no.fused_point_features(sd, feats, xyz[:, 2])
    pts = xyz.permute(0, 2, 3, 1).reshape(N, Hs * Ws, 3)
    mask = torch.all((pts < sd[no.PFX_C2M + "max_bound"]) & (pts >= sd[no.PFX_C2M + "min_bound"]), dim=2, keepdim=True)
    f = f * mask.view(N, Hs, Ws, 1).permute(0, 3, 1, 2)
    hom = torch.cat([pts, torch.ones_like(pts[:, :, :1])], dim=2)
    hom = (sd[no.PFX_C2M + "lidar2map"] @ hom.permute(0, 2, 1)).permute(0, 2, 1)
    xy = hom[:, :, :2] / sd[no.PFX_C2M + "voxel_size"][:2]
    vol, dens = splat_soft(xy, f.reshape(N, f.shape[1], Hs * Ws), grid[0], grid[1])
    return {"bev_features": vol.view(N, -1, grid[0], grid[1]), "bev_densities": dens.view(N, grid[0], grid[1], 1).permute(0, 3, 1, 2),
            "bev_coords": xy}


def vin_forward(sd, feat_map, ds=2, keys=("inpainting_sam_preds", "inpainting_sam_dynamic_preds", "elevation_preds")):
    iv = torch.cat([feat_map[k] for k in keys], dim=1)
    Ho, Wo = iv.shape[-2:]
    iv = F.max_pool2d(iv, ds, ds)
    iv = iv[:, :, : iv.shape[2] // 2, :]
    r = no.reward_fcn(sd, iv)
    full = torch.zeros(iv.shape[0], 1, Ho, Wo, device=iv.device)
    full[:, :, : Ho // 2, :] = F.interpolate(r, size=(Ho // 2, Wo), mode="bilinear", align_corners=False)
    return {"traversability_preds": r, "traversability_preds_full": full, "input_view": iv}


@torch.no_grad()
def forward(sd, rgbd, p2p):
    """rgbd [B,1,4,H,W], p2p [B,1,4,4] on any device -> the reference's output dict (solve_mdp=False)."""
    B, V, C, H, W = rgbd.shape
    out = depth_completion(sd, rgbd.view(B, C, H, W), (H, W))
    out["dino_pe_feats"] = no.dino_head(sd, out["depth_preds_feats"]).unsqueeze(1)
    out.update(cam2map(sd, out["depth_preds_metric"], out["depth_preds_feats"], p2p.view(B, 4, 4)))
    out.update(no.bev_decoder(sd, out["bev_features"]))
    out.update(vin_forward(sd, out))
    return out
