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


# ---------------------------------------------------------------------------------------------------------------
# On-device input pipeline.  The reference builds these warps on the CPU dataloader
# workers through kornia; here the per-pixel work is one CUDA kernel each (csrc/augment.cu) and only the 2x3 matrix
# algebra stays on the host, written out in the order kornia evaluates it (kornia is not a dependency of the mirror).
def _rotation_matrix2d(center, angle_deg, scale):
    """kornia.geometry.transform.get_rotation_matrix2d: center [1,2], angle [1] degrees, scale [1,2] -> [1,2,3]."""
    a = torch.deg2rad(angle_deg)
    c, s_ = torch.cos(a), torch.sin(a)
    rot = torch.stack([c, s_, -s_, c], dim=-1).view(-1, 2, 2)
    sr = rot @ torch.diag_embed(scale)
    alpha, beta = sr[:, 0, 0], sr[:, 0, 1]
    x, y = center[..., 0], center[..., 1]
    M = torch.zeros(center.shape[0], 2, 3, dtype=center.dtype)
    M[..., 0:2, 0:2] = sr
    M[..., 0, 2] = (1.0 - alpha) * x - beta * y
    M[..., 1, 2] = beta * x + (1.0 - alpha) * y
    return M


def _homography(M):
    Hm = F.pad(M, [0, 0, 0, 1], "constant", value=0.0)
    Hm[..., -1, -1] += 1.0
    return Hm


def get_affine_matrix2d(translations, center, scale, angle):
    """kornia.geometry.transform.get_affine_matrix2d (no shear): [B,3,3]."""
    t = _rotation_matrix2d(center, -angle, scale)
    t[..., 2] += translations
    return _homography(t)


def _pixel_to_norm(h, w, eps=1e-14):
    tr = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]])
    tr[0, 0] = tr[0, 0] * 2.0 / (eps if w == 1 else w - 1.0)
    tr[1, 1] = tr[1, 1] * 2.0 / (eps if h == 1 else h - 1.0)
    return tr.unsqueeze(0)


def affine_theta(M, src_hw, dst_hw):
    """The [B,2,3] theta kornia's warp_affine passes to F.affine_grid for the pixel-space map M [B,2,3] (src -> dst):
    normalise with the (size - 1) pixel transforms on both sides, invert."""
    M = M.detach().float().cpu()
    n_src, n_dst = _pixel_to_norm(*src_hw), _pixel_to_norm(*dst_hw)
    dst_norm_trans_src_norm = n_dst @ (_homography(M) @ torch.linalg.inv(n_src))
    return torch.linalg.inv(dst_norm_trans_src_norm)[:, :2, :]


def warp(input_tensor, transform, interpolation, precision=None, output_size=None, padding_mode="zeros"):
    """Reference creste/utils/utils.py:6-38 on the device: kornia warp_affine (align_corners=False) of the tensor plus a
    ones-channel; returns (output in the input's dtype, mask = warped ones > 0.99)."""
    if padding_mode != "zeros":
        raise NotImplementedError("warp: only zeros padding is used by the reference")
    assert input_tensor.ndim == 4 and transform.ndim == 3
    from creste_public_b200 import ops
    H, W = input_tensor.shape[-2:]
    out_hw = (H, W) if output_size is None else tuple(output_size)
    theta = affine_theta(transform, (H, W), out_hw)
    out, mask = ops.affine_warp(input_tensor.float(), theta, out_hw, nearest=(interpolation == "nearest"),
                                align_corners=False)
    return out.to(input_tensor.dtype), mask


class DepthAugmentation(object):
    """Reference train_utils.py:110-181 with the per-pixel work on the device (one fused pass: dropout mask, bilinear
    miscalibration warp, additive noise).  The random draws are made in the reference's order -- rand_like(depth),
    normal(calib mean, calib std), randn_like(depth) -- on the depth map's device (`draws` lets a caller supply them)."""

    def __init__(self, dropout_prob=0.1, calib_error_mean=[0.0, 0.0, 0.0], calib_error_std=[0.02, 0.02, 0.01],
                 depth_noise_std=0.2):
        self.dropout_prob = dropout_prob
        self.calib_error_mean = calib_error_mean
        self.calib_error_std = calib_error_std
        self.depth_noise_std = depth_noise_std

    def miscalibration_theta(self, noise, H, W):
        tx, ty = noise[0], noise[1]
        angle = noise[2] * (180.0 / torch.pi)
        T = get_affine_matrix2d(torch.tensor([[tx, ty]]), torch.tensor([[W / 2, H / 2]]), torch.tensor([[1.0, 1.0]]),
                                torch.tensor([angle]))
        return affine_theta(T[:, :2, :], (H, W), (H, W))[0]

    def __call__(self, depth_map, draws=None):
        assert depth_map.ndim == 3 and depth_map.shape[0] == 1, "Input depth map must have shape (1, H, W)."
        from creste_public_b200 import ops
        _, H, W = depth_map.shape
        if draws is None:
            u = torch.rand_like(depth_map)
            noise = torch.normal(mean=torch.tensor(self.calib_error_mean), std=torch.tensor(self.calib_error_std))
            g = torch.randn_like(depth_map)
        else:
            u, noise, g = draws
        theta = self.miscalibration_theta(noise.float().cpu(), H, W)
        return ops.depth_augment(depth_map, u, g, theta, self.dropout_prob, self.depth_noise_std)


class RotateAndTranslate(object):
    """Reference train_utils.py:183-318: the BEV map / FOV-mask warps run on the device (`transform_map`); the SE(2)
    matrices are host scalars.  `renew_transformation` draws like the reference."""

    def __init__(self, augmentations, map_size, voxel_size):
        self.augmentations = {}
        for aug in augmentations:
            kwargs = dict(aug)
            name = kwargs.pop("name")
            if name == "rotate":
                self.augmentations["rotate"] = kwargs.pop("max_rotation", 0.0)
            elif name == "translate":
                self.augmentations["translate"] = kwargs.pop("max_translation", 0.0)
            else:
                raise ValueError(f"Augmentation {name} not supported")
        self.map_size = torch.tensor(map_size).float()
        self.voxel_size = torch.tensor(voxel_size).float()
        self.center = (self.map_size / self.voxel_size / 2).float().unsqueeze(0)
        self.scale = torch.tensor([1.0, 1.0]).float().unsqueeze(0)
        assert len(self.augmentations) > 0

    def transform_map(self, map, R_init=None, interpolation="nearest"):
        """map [H,W,C] -> (tmap [H,W,C], mask [H,W])."""
        m = map.permute(2, 0, 1).unsqueeze(0)
        RT = self.mapRT
        if R_init is not None:
            RT = RT @ R_init
        tmap, mask = warp(m, RT, interpolation=interpolation)
        return tmap.squeeze(0).permute(1, 2, 0), mask.squeeze(0)

    def transform(self, inputs):
        assert inputs.shape[0] == 4, "Points must be Nx4 tensor"
        return self.RT.to(inputs.device) @ inputs

    def renew_transformation(self):
        RT = torch.eye(4)
        angle = torch.zeros(())
        if "translate" in self.augmentations:
            RT[:2, 3] = (2 * torch.rand(2) - 1) * self.augmentations["translate"]
        if "rotate" in self.augmentations:
            angle = (2 * torch.rand(1)[0] - 1) * self.augmentations["rotate"]
            a = angle * torch.pi / 180
            RT[:2, :2] = torch.tensor([[torch.cos(a), -torch.sin(a)], [torch.sin(a), torch.cos(a)]])
        self.RT = RT
        self.mapRT = _rotation_matrix2d(self.center, angle.reshape(1), self.scale)

    def compute_transformation_fromSE3(self, RT):
        R, t = RT[:2, :2], RT[:2, 3]
        offset = (t / self.voxel_size).float().unsqueeze(0)
        angle = (torch.atan2(R[1, 0], R[0, 0]) * 180 / torch.pi).reshape(1)
        return get_affine_matrix2d(offset, self.center, self.scale, angle)


def load_fov_mask(frustrum_mask, pc_augmentation, pose):
    """CodaPEFreeDataset._load_fov_mask (codapefree_dataloader.py:691-709; only the current pose contributes there):
    the trapezoidal frustum mask warped into the frame of `pose` on the device."""
    mask = frustrum_mask.clone().unsqueeze(-1).long()
    RT = pc_augmentation.compute_transformation_fromSE3(pose.cpu())
    mask, _ = pc_augmentation.transform_map(mask, R_init=RT)
    return mask.squeeze().bool()


def load_traverse(lidar_poses, voxel_size, bev_size):
    """CodaPEFreeDataset._load_traverse (codapefree_dataloader.py:590-615) from the relative LiDAR poses, on the device."""
    from creste_public_b200 import ops
    return ops.traverse_to_bev(lidar_poses, voxel_size, bev_size)


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
