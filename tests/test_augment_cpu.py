"""Host side of the on-device input pipeline: the kornia matrix algebra the mirror writes out
(creste/utils/train_utils.py) against the oracle's restatement, and the oracle's own invariants.  kornia is absent from
this image: the restated algorithm is UNPINNED by the reference (oracle/augment_oracle.py header); the sampling it
feeds is torch's own F.affine_grid / F.grid_sample."""
import math

import pytest
import torch

from oracle import augment_oracle as ao


def _tu():
    from creste_public_b200.creste.utils import train_utils
    return train_utils


def test_oracle_warp_affine_invariants():
    torch.manual_seed(0)
    x = torch.rand(1, 2, 9, 13)
    eye = torch.tensor([[[1.0, 0, 0], [0, 1, 0]]])
    assert float((ao.warp_affine(x, eye, (9, 13)) - x).abs().max()) <= 1e-6
    # an integer pixel translation moves the content exactly (align_corners=True, kornia's default)
    M = torch.tensor([[[1.0, 0, 3], [0, 1, 2]]])
    y = ao.warp_affine(x, M, (9, 13))
    assert float((y[:, :, 2:, 3:] - x[:, :, :-2, :-3]).abs().max()) <= 1e-5
    assert float(y[:, :, :2].abs().max()) <= 1e-6 and float(y[:, :, :, :3].abs().max()) <= 1e-6
    # rotation by 90 degrees about the centre of a square map is a transpose + flip
    s = torch.rand(1, 1, 8, 8)
    R = ao.get_rotation_matrix2d(torch.tensor([[3.5, 3.5]]), torch.tensor([90.0]), torch.ones(1, 2))
    r = ao.warp_affine(s, R, (8, 8))
    assert float((r - torch.rot90(s, 1, (2, 3))).abs().max()) <= 1e-5


def test_mirror_matrix_algebra_equals_oracle():
    tu = _tu()
    torch.manual_seed(1)
    for _ in range(5):
        c = torch.rand(1, 2) * 100
        ang = (torch.rand(1) - 0.5) * 90
        sc = torch.ones(1, 2)
        tr = torch.randn(1, 2) * 3
        assert torch.equal(tu.get_affine_matrix2d(tr, c, sc, ang), ao.get_affine_matrix2d(tr, c, sc, ang))
        M = ao.get_affine_matrix2d(tr, c, sc, ang)[:, :2, :]
        assert torch.equal(tu.affine_theta(M, (64, 96), (64, 96)), ao.warp_theta(M, (64, 96), (64, 96)))
        assert torch.equal(tu.affine_theta(M, (64, 96), (32, 48)), ao.warp_theta(M, (64, 96), (32, 48)))


def test_rotate_and_translate_matrices():
    tu = _tu()
    rt = tu.RotateAndTranslate([{"name": "rotate", "max_rotation": 20.0}, {"name": "translate", "max_translation": 2.0}],
                               [25.6, 25.6], [0.1, 0.1])
    assert torch.equal(rt.center, torch.tensor([[128.0, 128.0]]))
    torch.manual_seed(3)
    rt.renew_transformation()
    assert rt.RT.shape == (4, 4) and rt.mapRT.shape == (1, 2, 3)
    a = math.radians(17.0)
    pose = torch.eye(4)
    pose[:2, :2] = torch.tensor([[math.cos(a), -math.sin(a)], [math.sin(a), math.cos(a)]])
    pose[:2, 3] = torch.tensor([1.5, -0.7])
    M = rt.compute_transformation_fromSE3(pose)
    assert torch.equal(M, ao.se3_to_map_matrix(pose, rt.voxel_size, rt.center, rt.scale))
    with pytest.raises(ValueError):
        tu.RotateAndTranslate([{"name": "shear"}], [25.6, 25.6], [0.1, 0.1])


def test_oracle_load_traverse_clamps_to_the_grid():
    P = torch.eye(4).repeat(4, 1, 1)
    P[:, 0, 3] = torch.tensor([0.0, 5.0, 20.0, -20.0])
    P[:, 1, 3] = torch.tensor([0.0, -3.0, 1.0, 2.0])
    G = ao.load_traverse(P, torch.tensor([0.1, 0.1]), (256, 256))
    assert torch.equal(G[0, :2, 2], torch.tensor([128.0, 128.0]))
    assert torch.equal(G[1, :2, 2], torch.tensor([78.0, 158.0]))
    assert torch.equal(G[2, :2, 2], torch.tensor([0.0, 118.0]))        # clamped at the far edge
    assert torch.equal(G[3, :2, 2], torch.tensor([256.0, 108.0]))      # clamped at the near edge
