"""Seeded synthetic inputs for the tests / oracle: re-exports the pure generators of the top-level
`synth_data` module and adds the one generator that runs oracle code (the LiDAR raster of the
network input).  TEST INFRASTRUCTURE."""
import torch

from synth_data import *  # noqa: F401,F403
from synth_data import (FIXED_BUFFERS, T_CAM_TO_LIDAR, _rng, lidar2camrect, make_p2p, os1_scan,  # noqa: F401
                        rgb_frames)


def net_inputs(H, W, B=1, seed=0):
    """Synthetic RGB + rasterised LiDAR depth (mm) + p2p, SURVEY.md section 8(d) config 2."""
    from . import c_oracle
    rgb = torch.from_numpy(rgb_frames(B, H, W, seed))
    frames = []
    for b in range(B):
        _, dmm = c_oracle.lidar_raster(os1_scan(seed + b), lidar2camrect(H, W), H, W)
        frames.append(torch.from_numpy(dmm).view(1, 1, 1, H, W))
    rgbd = torch.cat([rgb, torch.cat(frames, 0)], dim=2)
    p2p = torch.from_numpy(make_p2p(H, W)).view(1, 1, 4, 4).repeat(B, 1, 1, 1)
    return rgbd, p2p
