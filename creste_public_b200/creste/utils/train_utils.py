"""Hot-path helpers of reference creste/utils/train_utils.py: create_trapezoidal_fov_mask
(:511-557), prefix_dict / merge_dict / merge_loss_dict (:560-599), resize_and_crop (:670-682).
The FOV mask is an init-time constant (a few KB) built once on the host, as in the reference."""
import math

import torch
import torch.nn.functional as F


def create_trapezoidal_fov_mask(H, W, fov_top_angle=50, fov_bottom_angle=40, near=10, far=50):
    """Boolean [H,W] mask of a north-facing trapezoidal field of view."""
    ys = torch.arange(H).view(H, 1).expand(H, W)
    xs = torch.arange(W).view(1, W).expand(H, W)
    dx, dy = xs - W / 2, ys - H / 2
    dist = torch.sqrt(dx ** 2 + dy ** 2)
    # north is "up": the second argument is (H/2 - y), NOT -(y - H/2) -- at the centre cell the former is
    # +0.0 (angle 0, inside), the latter -0.0 (angle 180)
    ang = torch.atan2(dx, H / 2 - ys) * 180 / torch.pi
    ang = torch.where(ang < -180, ang + 360, ang)
    top = torch.full_like(dist, fov_top_angle / 2)
    bot = torch.full_like(dist, fov_bottom_angle / 2)
    mid = top + (bot - top) * ((dist - near) / (far - near))
    spread = torch.where(dist <= near, top, torch.where(dist >= far, bot, mid))
    return (dist >= near) & (dist <= far) & (ang.abs() <= spread)


def prefix_dict(prefix, d, seprator="/"):
    return {prefix + seprator + k: v for k, v in d.items()}


def merge_dict(*args):
    ret = {}
    for arg in args:
        if isinstance(arg, dict):
            ret.update(arg)
        else:
            ret.update(prefix_dict(arg[0], arg[1]))
    return ret


def merge_loss_dict(full_dict, new_dict):
    full_dict.update(new_dict)
    return full_dict


def resize_and_crop(x, size, crop):
    """Nearest resize to `size` then crop (top, bottom, left, right) -- used on tiny masks."""
    x = F.interpolate(x, size=size, mode="nearest")
    t, b, l, r = crop
    return x[..., t:b, l:r]


def earliest_pose_in_fov(expert, fov_mask, return_idx=False):
    """First expert pose that lies inside the FOV mask (reference train_utils.py:765-803).
    expert [B,T,2] (row, col; truncated to integers), fov_mask [1,1,H,W] -> S0 [B,2] int64; samples with no
    pose in view start at the bottom-centre cell (H-1, W//2).  With return_idx also the earliest index
    (0 when none) and the latest index (-1 when none).  The stage-3 hot path has this rule fused into
    creste_svf; this stand-alone form is a handful of index ops on a [B,T] tensor."""
    B, T, _ = expert.shape
    H, W = fov_mask.shape[-2:]
    rows, cols = expert[..., 0].long(), expert[..., 1].long()
    inside = fov_mask.to(rows.device)[0, 0][rows, cols] == 1                      # [B,T]
    t = torch.arange(T, device=rows.device).expand(B, T)
    first = torch.where(inside, t, torch.full_like(t, T)).amin(dim=1)
    last = torch.where(inside, t, torch.full_like(t, -1)).amax(dim=1)
    none = first == T
    first = torch.where(none, torch.zeros_like(first), first)
    pick = first.view(B, 1)
    pose = torch.cat([rows.gather(1, pick), cols.gather(1, pick)], dim=1)
    default = torch.tensor([H - 1, W // 2], dtype=pose.dtype, device=pose.device)
    pose = torch.where(none.view(B, 1), default.view(1, 2), pose)
    return (pose, first, last) if return_idx else pose


def gaussian_2d(goals, sigma, H, W):
    """Isotropic Gaussian bumps [B,1,H,W] centred at goals [B,2] = (row, col) (reference :806-834; used by the
    disabled goal_kwargs branch only)."""
    B = goals.size(0)
    r = torch.arange(H, device=goals.device).float().view(1, H, 1) - goals[:, 0].float().view(B, 1, 1)
    c = torch.arange(W, device=goals.device).float().view(1, 1, W) - goals[:, 1].float().view(B, 1, 1)
    return torch.exp(-(r ** 2 + c ** 2) / (2 * sigma ** 2)).view(B, 1, H, W)


def get_save_paths(cfg, model_type="ssc", stage="train"):
    """Checkpoint directory of a run (reference :602-667): <root>/<project>/<run name>/<day>/<time>, created
    on demand.  The run name is assembled from the backbone / head / optimiser entries of the config."""
    import os
    from datetime import datetime
    day, clock = datetime.now().strftime("%Y%m%d_%H%M%S").split("_")
    model_cfg, suffix = cfg["model"], ""
    if model_type == "lfd_maxentirl":
        suffix = "/%s_head%s_horizon%s" % (cfg["model"]["run_name"], cfg["model"]["traversability_head"]["name"],
                                           cfg["dataset"]["action_horizon"])
        model_cfg = cfg["model"]["vision_backbone"]
    elif model_type == "lfd_bc":
        suffix = "/%s_in%s_out%s_lr%s" % (cfg["model"]["run_name"], cfg["model"]["bc_head"]["in_horizon"],
                                          cfg["model"]["bc_head"]["out_horizon"], cfg["model"]["optimizer"]["lr"])
        model_cfg = cfg["model"]["backbone"]
    run_name = suffix
    if model_type != "cluster_probe":
        run_name = "%s_BB_%s_Head_%s_lr_%f_%s_%s_v2" % (
            model_cfg["run_name"], model_cfg["vision_backbone"]["name"], model_cfg["depth_head"]["name"],
            model_cfg["optimizer"]["lr"], model_cfg["discretize"]["mode"], cfg["dataset"]["infill_strat"]) + suffix
    if stage not in ("train", "test"):
        raise ValueError(f"Invalid stage {stage}")
    out = os.path.join(cfg["trainer"]["default_root_dir"], cfg["model"]["project_name"], run_name, day, clock)
    os.makedirs(out, exist_ok=True)
    return out


def extract_max_per_class(tensor, max_per_class=100, return_indices=True):
    """Up to `max_per_class` random members of every class of a 1-D label tensor (reference :324-352).  The random
    subsets are drawn with torch.randperm from the DEFAULT (CPU) generator, class by class in ascending label
    order -- the same draws as the reference's on the same seed."""
    picked = []
    for cls in torch.unique(tensor):
        idx = (tensor == cls).nonzero(as_tuple=False).reshape(-1)
        if idx.size(0) > max_per_class:
            idx = idx[torch.randperm(idx.size(0))[:max_per_class].to(idx.device)]
        picked.append(idx)
    out = torch.cat(picked) if picked else torch.zeros(0, dtype=torch.long, device=tensor.device)
    return out if return_indices else tensor[out]


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "utils/train_utils.py")
