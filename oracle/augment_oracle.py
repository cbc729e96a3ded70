"""TEST INFRASTRUCTURE (oracle): the reference's input-pipeline warps restated on plain torch CPU.

Parity status: UNPINNED BY THE REFERENCE for the kornia part.  `kornia` (un-pinned in the reference's requirements; the
algorithm below is that of kornia 0.7.x `kornia/geometry/transform/imgwarp.py` and `kornia/geometry/conversions.py`) is
absent from this image and from /root/reference, so its published algorithm is restated here:
  get_rotation_matrix2d / get_affine_matrix2d / normal_transform_pixel / normalize_homography / warp_affine
(warp_affine = convert to 3x3, normalise with the (size - 1) pixel transform on both sides, invert, F.affine_grid +
F.grid_sample).  The sampling itself IS pinned: it is torch's own F.affine_grid / F.grid_sample, the functions kornia
calls.  On top of that, restated from the reference's own files:
  warp              creste/utils/utils.py:6-38
  depth_augment     creste/utils/train_utils.py:110-181 (DepthAugmentation.__call__, given the random draws)
  se3_to_map_matrix creste/utils/train_utils.py:301-318 (RotateAndTranslate.compute_transformation_fromSE3)
  load_traverse     creste/datasets/codapefree_dataloader.py:579-615 (given the relative LiDAR poses)
Only tests/ may import this module."""
import math

import torch
import torch.nn.functional as F


def angle_to_rotation_matrix(angle):
    a = angle * math.pi / 180.0 if not torch.is_tensor(angle) else torch.deg2rad(angle)
    c, s = torch.cos(a), torch.sin(a)
    return torch.stack([c, s, -s, c], dim=-1).view(*a.shape, 2, 2)


def get_rotation_matrix2d(center, angle, scale):
    rot = angle_to_rotation_matrix(angle)                       # [B,2,2]
    scaling = torch.zeros(center.shape[0], 2, 2, dtype=center.dtype)
    scaling[:, 0, 0] = scale[:, 0]
    scaling[:, 1, 1] = scale[:, 1]
    sr = rot @ scaling
    alpha, beta = sr[:, 0, 0], sr[:, 0, 1]
    x, y = center[..., 0], center[..., 1]
    M = torch.zeros(center.shape[0], 2, 3, dtype=center.dtype)
    M[..., 0:2, 0:2] = sr
    M[..., 0, 2] = (1.0 - alpha) * x - beta * y
    M[..., 1, 2] = beta * x + (1.0 - alpha) * y
    return M


def to_h(M):
    H = F.pad(M, [0, 0, 0, 1], "constant", value=0.0)
    H[..., -1, -1] += 1.0
    return H


def get_affine_matrix2d(translations, center, scale, angle):
    t = get_rotation_matrix2d(center, -angle, scale)
    t[..., 2] += translations
    return to_h(t)


def normal_transform_pixel(h, w, eps=1e-14):
    tr = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]])
    wd = eps if w == 1 else w - 1.0
    hd = eps if h == 1 else h - 1.0
    tr[0, 0] = tr[0, 0] * 2.0 / wd
    tr[1, 1] = tr[1, 1] * 2.0 / hd
    return tr.unsqueeze(0)


def normalize_homography(dst_pix_trans_src_pix, src_hw, dst_hw):
    src_norm_trans_src_pix = normal_transform_pixel(*src_hw).to(dst_pix_trans_src_pix)
    src_pix_trans_src_norm = torch.linalg.inv(src_norm_trans_src_pix)
    dst_norm_trans_dst_pix = normal_transform_pixel(*dst_hw).to(dst_pix_trans_src_pix)
    return dst_norm_trans_dst_pix @ (dst_pix_trans_src_pix @ src_pix_trans_src_norm)


def warp_theta(M, src_hw, dsize):
    """The [B,2,3] theta kornia hands to F.affine_grid for the pixel-space map M ([B,2,3], src -> dst)."""
    dst_norm_trans_src_norm = normalize_homography(to_h(M), src_hw, dsize)
    return torch.linalg.inv(dst_norm_trans_src_norm)[:, :2, :]


def warp_affine(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=True):
    B, C, H, W = src.shape
    theta = warp_theta(M, (H, W), dsize)
    grid = F.affine_grid(theta, [B, C, dsize[0], dsize[1]], align_corners=align_corners)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


def warp(input_tensor, transform, interpolation, output_size=None, padding_mode="zeros"):
    inp = input_tensor.to(transform.dtype)
    inp_plus_mask = F.pad(inp, (0, 0, 0, 0, 0, 1), value=1.0)
    if output_size is None:
        output_size = inp_plus_mask.shape[-2:]
    w = warp_affine(inp_plus_mask, transform, tuple(output_size), mode=interpolation, padding_mode=padding_mode,
                    align_corners=False)
    return w[:, :-1].to(input_tensor.dtype), w[:, -1] > 0.99


def depth_augment(depth, u, calib_noise, g, dropout_prob=0.1, depth_noise_std=0.2):
    """DepthAugmentation.__call__ on a [1,H,W] map given its draws: u = rand_like(depth), calib_noise = the three
    normal draws (tx, ty, angle in radians), g = randn_like(depth)."""
    _, H, W = depth.shape
    d = depth * (u > dropout_prob)
    tx, ty = calib_noise[0], calib_noise[1]
    angle = calib_noise[2] * (180.0 / torch.pi)
    center = torch.tensor([[W / 2, H / 2]])
    T = get_affine_matrix2d(torch.tensor([[tx, ty]]), center, torch.tensor([[1.0, 1.0]]), torch.tensor([angle]))
    d = warp_affine(d.unsqueeze(0), T[:, :2, :], (H, W)).squeeze(0)
    return d + g * depth_noise_std


def se3_to_map_matrix(RT, voxel_size, center, scale):
    R, t = RT[:2, :2], RT[:2, 3]
    offset = (t / voxel_size).float().unsqueeze(0)
    angle = (torch.atan2(R[1, 0], R[0, 0]) * 180 / torch.pi).reshape(1)
    return get_affine_matrix2d(offset, center, scale, angle)


def load_traverse(lidar_poses, voxel_size, bev_size):
    T = lidar_poses.shape[0]
    P = torch.eye(3, 3).repeat(T, 1, 1)
    P[:, :2, :2] = lidar_poses[:, :2, :2]
    P[:, :2, 2] = lidar_poses[:, :2, 3] / voxel_size
    T_lidar_to_bev = torch.tensor([[-1, 0, bev_size[1] // 2], [0, -1, bev_size[0] // 2], [0, 0, 1]], dtype=torch.float32)
    G = torch.matmul(T_lidar_to_bev, P)
    G[:, :2, 2] = torch.clamp(G[:, :2, 2], torch.tensor([0, 0]).float(), torch.tensor(bev_size).float())
    return G
