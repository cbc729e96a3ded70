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
    ang = torch.atan2(dx, -dy) * 180 / torch.pi
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
