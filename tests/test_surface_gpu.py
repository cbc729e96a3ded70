"""GPU tests of the reference's helper methods exposed as callable entry points on the mirror (Camera2World.forward,
Camera2MapMulti._points_to_voxels / splat_soft / _prepare_features_and_coords, creste.utils.depth_utils) and of the
pack-cache invalidation after in-process training (eval -> train step -> eval)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import net_oracle as no
from oracle import synth

pytestmark = pytest.mark.gpu


def _terrainnet(hw=(64, 96)):
    import creste_public_b200 as cb
    cb.set_precision("fp32")
    m = cb.build_terrainnet(image_size=hw).eval()
    sd = synth.seeded_state_dict(m.state_dict(), 0, "peaky")
    m.load_state_dict(sd)
    return m.cuda(), sd


def test_camera2world_forward_matches_reference_math(cuda):
    """reference splat_projection.py:19-51 restated on torch CPU (meshgrid, [u*d, v*d, d, 1], bmm)."""
    m, _ = _terrainnet()
    B, N, Hs, Ws = 2, 1, 16, 24
    g = torch.Generator().manual_seed(0)
    depth = torch.rand(B, N, Hs, Ws, generator=g) * 20 + 0.3
    p2p = torch.from_numpy(synth.make_p2p(64, 96)).view(1, 1, 4, 4).repeat(B, N, 1, 1)
    u, v = torch.meshgrid(torch.arange(Ws), torch.arange(Hs), indexing="xy")
    cam = torch.stack([u, v, torch.ones_like(u)], 0).unsqueeze(0).repeat(B * N, 1, 1, 1) * depth.view(B * N, 1, Hs, Ws)
    cam = torch.cat([cam, torch.ones(B * N, 1, Hs, Ws)], 1)
    want = torch.bmm(p2p.view(B * N, 4, 4), cam.flatten(2)).view(B, N, 4, Hs, Ws)[:, :, :3]
    got = m.cam2map.cam2world((depth.cuda(), p2p.cuda()))
    assert tuple(got.shape) == (B, N, 3, Hs, Ws)
    assert torch.equal(got.cpu(), want)


def test_points_to_voxels_and_splat_soft_methods(cuda):
    m, _ = _terrainnet()
    c2m = m.cam2map
    g = torch.Generator().manual_seed(1)
    pts = (torch.rand(2, 500, 3, generator=g) - 0.5) * 30
    want = torch.cat([pts, torch.ones(2, 500, 1)], 2)
    want = (c2m.lidar2map.cpu() @ want.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :2] / c2m.voxel_size.cpu()[:2]
    got = c2m._points_to_voxels(pts.cuda())
    assert torch.equal(got.cpu(), want)
    # splat_soft((xy, feats [B,F,P], grid)) -> (features [B,F,G], densities [B,G,1]) vs the C oracle
    depth, p2p, feats = synth.splat_inputs(seed=5, N=2, Hs=16, Ws=24, F=8)          # feats [N,F,P]
    rng = [float(v) for v in c2m.point_cloud_range.tolist()]
    _, xy, _ = co.frustum_to_bev(depth, p2p, rng, [float(v) for v in c2m.voxel_size.tolist()][:2])
    vf, vd = c2m.splat_soft((torch.from_numpy(xy).cuda(), torch.from_numpy(feats).cuda(), c2m.grid_size[:2]))
    bev, dens, _, _ = co.splat_soft(xy, feats, 256, 256)
    assert tuple(vf.shape) == (2, 8, 256 * 256) and tuple(vd.shape) == (2, 256 * 256, 1)
    np.testing.assert_allclose(vf.cpu().numpy().reshape(bev.shape), bev, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(vd.cpu().numpy().reshape(dens.shape), dens, atol=1e-5, rtol=1e-5)


def test_prepare_features_and_coords_method(cuda):
    import creste_public_b200 as cb
    cb.set_precision("fp32")
    irl = cb.build_maxentirl(image_size=(64, 96)).eval()
    sd = synth.seeded_state_dict(irl.state_dict(), 0, "peaky")
    irl.load_state_dict(sd)
    m = irl.cuda().backbone
    rgbd, p2p = synth.net_inputs(64, 96, 1)
    ref = no.forward(sd, rgbd, p2p)
    depth, feats = ref["depth_preds_metric"], ref["depth_preds_feats"]
    B, F, Hs, Ws = feats.shape
    with torch.no_grad():
        xyz, mask, fused = m.cam2map._prepare_features_and_coords(
            (depth.view(B, 1, Hs, Ws).cuda(), feats.view(B, 1, F, Hs, Ws).cuda(), p2p.cuda()))
    assert tuple(xyz.shape) == (B, 1, 3, Hs, Ws) and mask.dtype == torch.bool
    assert tuple(mask.shape) == (B, 1, 1, Hs, Ws) and tuple(fused.shape) == (B, 1, 96, Hs, Ws)
    r = ref["_fused_feats"]              # the oracle keeps the bounds-masked features (what the splat consumes)
    got = fused.cpu().view_as(r) * mask.cpu().view(B, 1, Hs, Ws).float()
    assert float((got - r).abs().max()) <= 1e-5 * float(r.abs().max())
    lo, hi = m.cam2map.min_bound.cpu().view(1, 1, 3, 1, 1), m.cam2map.max_bound.cpu().view(1, 1, 3, 1, 1)
    want_mask = ((xyz.cpu() < hi) & (xyz.cpu() >= lo)).all(dim=2, keepdim=True)
    assert torch.equal(mask.cpu(), want_mask)


def test_depth_utils_mirror(cuda):
    from creste_public_b200.creste.utils import depth_utils as du
    g = torch.Generator().manual_seed(2)
    d = torch.rand(2, 1, 16, 24, generator=g) * 30000 - 1000
    d[0, 0, 0, 0] = float("inf")
    d[0, 0, 0, 1] = float("nan")
    bs = (25600 - 300) / 128
    idx = (d - 300) / bs
    want_f = idx.clone()
    bad = (idx < 0) | (idx > 128) | ~torch.isfinite(idx)
    want_i = torch.where(bad, torch.full_like(idx, 128.0), idx).nan_to_num(128.0).long()
    want_i[bad] = 128
    got_f = du.bin_depths(d.cuda(), "UD", 300, 25600, 128, target=False).cpu()
    got_i = du.bin_depths(d.cuda(), "UD", 300, 25600, 128, target=True).cpu()
    ok = torch.isfinite(want_f)
    assert torch.equal(got_f[ok], want_f[ok]) and got_i.dtype == torch.int64 and torch.equal(got_i, want_i)
    lid = du.bin_depths(d.clamp(300, 25600).cuda(), "LID", 300, 25600, 128).cpu()
    bs_l = 2 * (25600 - 300) / (128 * 129)
    np.testing.assert_allclose(lid.numpy(), (-0.5 + 0.5 * torch.sqrt(1 + 8 * (d.clamp(300, 25600) - 300) / bs_l)).numpy(),
                               rtol=1e-6, atol=1e-5)
    nchw = torch.from_numpy(synth.depth_logits_inputs())             # [2, 128, 6, 10]
    want = (torch.softmax(nchw, 1) * torch.linspace(300, 25600, 128).view(1, -1, 1, 1)).sum(1)
    got = du.convert_to_metric_depth_differentiable(nchw.cuda(), "UD", 300, 25600, 128).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=3e-6)


def test_eval_after_training_step_uses_new_parameters(cuda):
    """eval forward -> stage-1 training step (fused Adam + BatchNorm running statistics written through raw
    pointers) -> eval forward: the second eval must see the updated parameters, i.e. equal a freshly built model
    loaded with the post-step state dict (the pack caches key on tensor versions)."""
    import creste_public_b200 as cb
    from creste_public_b200 import configs
    from creste_public_b200.creste.train_pefree import DistillationModel
    import synth_data
    cb.set_precision("fp32")
    hw = (64, 96)
    torch.manual_seed(0)
    dm = DistillationModel(configs.distill_cfg(hw)).cuda()
    batch = {k: v.cuda() for k, v in synth_data.distill_batch(2, hw[0], hw[1], seed=0).items()}
    dm.eval()
    with torch.no_grad():
        before = dm(batch["image"])["depth_preds_logits"].clone()
    dm.train()
    dm.optimizers().lr = 1e-2
    dm.training_step(batch)
    dm.eval()
    with torch.no_grad():
        after = dm(batch["image"])["depth_preds_logits"].clone()
    assert float((after - before).abs().max()) > 1e-3          # the step moved the weights
    fresh = DistillationModel(configs.distill_cfg(hw)).cuda().eval()
    fresh.model.load_state_dict(dm.model.state_dict())
    with torch.no_grad():
        want = fresh(batch["image"])["depth_preds_logits"]
    assert torch.equal(after, want)
