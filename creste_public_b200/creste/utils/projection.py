"""Mirror of the hot-path part of reference creste/utils/projection.py: the calibration
transforms `get_pixel2pts_transform` (:11-34) / `get_pts2pixel_transform` (:37-61) -- 4x4 float64
host matrices, built once per frame by the dataloader (codapefree_dataloader.py:803-841) -- and
`pixels_to_depth` (:64-155), the LiDAR -> sparse depth raster that feeds channel 3 of the RGB-D
input.  The raster runs on the GPU (`creste_lidar_raster`: float64 projection, truncation toward
zero, per-pixel maximum) instead of numpy + torch_scatter on the host.
"""
import numpy as np
import torch

from creste_public_b200 import ops


def _h(m3x4_or_4x4):
    out = np.eye(4)
    out[:3, :] = np.asarray(m3x4_or_4x4, dtype=np.float64)[:3, :]
    return out


def _rot(r3x3):
    out = np.eye(4)
    out[:3, :3] = np.asarray(r3x3, dtype=np.float64)
    return out


def get_pixel2pts_transform(calib_dict):
    """[4,4] float64: homogeneous pixel (u*d, v*d, d, 1) in the rectified image -> LiDAR frame,
    inv(T_lidar->cam) @ R^T @ inv(P[:3,:3])."""
    return np.linalg.inv(_h(calib_dict["lidar2cam"])) @ _rot(np.asarray(calib_dict["R"]).T) \
        @ _rot(np.linalg.inv(np.asarray(calib_dict["P"], dtype=np.float64)[:3, :3]))


def get_pts2pixel_transform(calib_dict):
    """[4,4] float64: LiDAR point -> rectified pixel, P[:3,:3] @ R @ T_lidar->cam."""
    return _rot(np.asarray(calib_dict["P"], dtype=np.float64)[:3, :3]) @ _rot(calib_dict["R"]) \
        @ _h(calib_dict["lidar2cam"])


def lidar_depth_raster(pc, lidar2camrect, IMG_H, IMG_W, device=None):
    """pc [N,>=3] (numpy or tensor), lidar2camrect [3+,4] -> (depth_m [H,W], depth_mm [H,W]) CUDA
    tensors; depth_mm = uint16-truncated millimetres stored as float32, the network's channel 3
    (scripts/preprocessing/build_dense_depth.py:461-463)."""
    if not isinstance(pc, torch.Tensor):
        pc = torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float32))
    if device is None:
        device = pc.device if pc.is_cuda else torch.device("cuda")
    pc = pc.to(device=device, dtype=torch.float32)
    P = lidar2camrect.cpu().numpy() if isinstance(lidar2camrect, torch.Tensor) else np.asarray(lidar2camrect)
    return ops.lidar_raster(pc, P[:3, :4].astype(np.float64), IMG_H, IMG_W)


def pixels_to_depth(pc_np, calib, IMG_H, IMG_W, return_keys=("image_pts", "image_depth"),
                    IMG_DEBUG_FLAG=False, depth_priority="max"):
    """Reference signature and default return values: `image_pts` [M,2] int (x, y) of the pixels
    that received a point, row-major, and `image_depth` [M] their farthest depth in metres."""
    if depth_priority != "max":
        raise NotImplementedError("depth_priority other than 'max' is unused by the reference pipeline")
    depth_m, _ = lidar_depth_raster(pc_np, calib["lidar2camrect"], IMG_H, IMG_W)
    img = depth_m.cpu().numpy()
    out = []
    for key in return_keys:
        if key == "image_pts":
            ys, xs = np.nonzero(img)
            out.append(np.stack([xs, ys], axis=1))
        elif key == "image_depth":
            out.append(img[img != 0].reshape(-1))
        else:
            raise NotImplementedError(f"return key {key!r}: only the rasterised outputs are on the hot path")
    return out


def make_rgbd(rgb, pc, lidar2camrect):
    """CodaPEFreeDataset._load_rgbd (codapefree_dataloader.py:843-879) on the device:
    rgb uint8 / float [3,H,W] (0..255) + LiDAR sweep -> float32 [4,H,W] = cat(rgb / 255, depth_mm)."""
    rgb = rgb if isinstance(rgb, torch.Tensor) else torch.from_numpy(np.asarray(rgb))
    rgb = rgb.cuda().float()
    _, H, W = rgb.shape
    out = torch.empty(4, H, W, device=rgb.device)
    out[:3] = rgb / 255.0
    ops.lidar_raster(pc if isinstance(pc, torch.Tensor) else torch.from_numpy(np.asarray(pc, np.float32)).cuda(),
                     np.asarray(lidar2camrect)[:3, :4].astype(np.float64), H, W, out_mm=out[3], want_m=False)
    return out


# names this mirror does not define fall through to the reference's file when the mirror is overlaid on a checkout
from creste_public_b200.creste import _overlay  # noqa: E402
__getattr__ = _overlay.fallback(__name__, "utils/projection.py")
