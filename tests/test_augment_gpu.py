"""GPU parity of the on-device input pipeline kernels (csrc/augment.cu) against oracle/augment_oracle.py: kornia's
warp_affine algebra restated (unpinned: kornia is absent) on torch's own F.affine_grid / F.grid_sample (pinned).
Bilinear outputs: 1e-4 of the tensor maximum (the sampling coordinates go through a float matmul whose summation order
differs); nearest outputs and validity masks: identical away from rounding ties -- at most 0.1 % of the pixels may
differ (a tie at x.5 is decided by the last bit of the coordinate)."""
import math

import pytest
import torch

from oracle import augment_oracle as ao

pytestmark = pytest.mark.gpu


def _pose(deg, tx, ty):
    a = math.radians(deg)
    P = torch.eye(4)
    P[:2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
    P[:2, 3] = torch.tensor([tx, ty])
    return P


@pytest.mark.parametrize("H,W,C,deg,tx,ty,mode", [(64, 96, 3, 11.0, 4.5, -2.25, "bilinear"), (64, 96, 1, -33.0, 0.0, 7.0, "nearest"),
                                                  (256, 256, 2, 5.0, 30.0, -12.0, "nearest"), (33, 17, 4, 90.0, 1.0, 1.0, "bilinear")])
def test_warp_matches_oracle(cuda, H, W, C, deg, tx, ty, mode):
    from creste_public_b200.creste.utils import train_utils as tu
    torch.manual_seed(H + C)
    x = torch.rand(1, C, H, W)
    M = ao.get_affine_matrix2d(torch.tensor([[tx, ty]]), torch.tensor([[W / 2, H / 2]]), torch.ones(1, 2),
                               torch.tensor([deg]))[:, :2, :]
    ref, ref_mask = ao.warp(x, M, mode)
    out, mask = tu.warp(x.to(cuda), M, mode)
    assert out.shape == ref.shape and mask.shape == ref_mask.shape and mask.dtype == torch.bool
    if mode == "bilinear":
        assert float((out.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    else:
        assert float((out.cpu() != ref).float().mean()) <= 1e-3
    assert float((mask.cpu() != ref_mask).float().mean()) <= 1e-3
    # smaller output canvas
    ref2, m2 = ao.warp(x, M, mode, output_size=(H // 2, W // 2))
    out2, mk2 = tu.warp(x.to(cuda), M, mode, output_size=(H // 2, W // 2))
    if mode == "bilinear":
        assert float((out2.cpu() - ref2).abs().max()) <= 1e-4 * max(1e-6, float(ref2.abs().max()))
    else:
        assert float((out2.cpu() != ref2).float().mean()) <= 2e-3
    assert float((mk2.cpu() != m2).float().mean()) <= 2e-3


def test_depth_augmentation_matches_oracle_given_the_draws(cuda):
    from creste_public_b200.creste.utils import train_utils as tu
    torch.manual_seed(5)
    H, W = 128, 240
    depth = torch.rand(1, H, W) * 20000.0 * (torch.rand(1, H, W) > 0.7)        # sparse LiDAR raster, mm
    aug = tu.DepthAugmentation()
    u = torch.rand_like(depth)
    noise = torch.normal(mean=torch.tensor(aug.calib_error_mean), std=torch.tensor(aug.calib_error_std)) * 20.0
    g = torch.randn_like(depth)
    ref = ao.depth_augment(depth, u, noise, g, aug.dropout_prob, aug.depth_noise_std)
    out = aug(depth.to(cuda), draws=(u.to(cuda), noise, g.to(cuda)))
    assert out.shape == ref.shape
    assert float((out.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    # own draws on the device: same statistics (dropout rate, noise level on empty pixels)
    torch.manual_seed(6)
    out2 = aug(torch.zeros(1, H, W, device=cuda))
    assert abs(float(out2.std()) - aug.depth_noise_std) <= 0.02 and abs(float(out2.mean())) <= 0.01


def test_fov_mask_warp_and_traverse(cuda):
    from creste_public_b200.creste.utils import train_utils as tu
    rt = tu.RotateAndTranslate([{"name": "rotate", "max_rotation": 0.0}, {"name": "translate", "max_translation": 0.0}],
                               [25.6, 25.6], [0.1, 0.1])
    torch.manual_seed(0)
    rt.renew_transformation()
    frustum = tu.create_trapezoidal_fov_mask(256, 256, 70, 70, 7, 200)
    pose = _pose(23.0, 1.5, -0.8)
    got = tu.load_fov_mask(frustum.to(cuda), rt, pose)
    M = rt.mapRT @ ao.se3_to_map_matrix(pose, rt.voxel_size, rt.center, rt.scale)
    ref, _ = ao.warp(frustum.clone().unsqueeze(-1).long().permute(2, 0, 1).unsqueeze(0), M, "nearest")
    ref = ref.squeeze().bool()
    assert got.shape == ref.shape and got.dtype == torch.bool
    assert float((got.cpu() != ref).float().mean()) <= 1e-3 and 0.02 < float(got.float().mean()) < 0.9
    # expert trajectory -> BEV grid poses: exact
    T = 50
    P = torch.stack([_pose(2.0 * t, 0.4 * t, 0.05 * t * (-1) ** t) for t in range(T)])
    ref_g = ao.load_traverse(P, torch.tensor([0.1, 0.1]), (256, 256))
    got_g = tu.load_traverse(P.to(cuda), [0.1, 0.1], (256, 256))
    assert torch.equal(got_g.cpu(), ref_g)
