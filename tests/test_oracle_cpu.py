"""CPU tests (-m "not gpu"): the oracle against the golden vectors minted from the unmodified
reference (tests/golden/*.npz, generator oracle/gen_golden.py), and -- when the reference tree is
present (build container only) -- against the reference's own code executed under the shims."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import net_oracle as no
from oracle import ref_shims, synth

HAVE_REF = ref_shims.reference_available()


@pytest.mark.parametrize("name", ["b2_16x16", "b1_64x64", "b3_24x40"])
def test_vi_oracle_matches_golden(golden, name):
    g = golden("vi.npz")
    seed, B, H, W, K = [int(x) for x in g[f"{name}_meta"]]
    v, q, pi, k = co.vi_solve(synth.vi_inputs(seed, B, H, W))
    assert k == K
    assert np.array_equal(v.view(np.uint32), g[f"{name}_v"][:, 0].view(np.uint32))
    if f"{name}_q" in g.files:
        assert np.array_equal(q.view(np.uint32), g[f"{name}_q"].view(np.uint32))
        np.testing.assert_allclose(pi, g[f"{name}_pi"], atol=1e-6)


@pytest.mark.parametrize("name", ["b2_32x64", "b2_32x64_zt", "b1_64x128"])
def test_svf_oracle_matches_golden(golden, name):
    g = golden("svf.npz")
    seed, B, H, W, T, zt = [int(x) for x in g[f"{name}_meta"]]
    r, expert = synth.svf_inputs(seed, B, H, W, T)
    v, q, pi, k = co.vi_solve(r)
    s, st, grid = co.svf(pi, expert[:, :, :2, 2].copy(), g[f"{name}_fov"], T, 2, True, 0.005, bool(zt))
    assert np.array_equal(st, g[f"{name}_states"]) and np.array_equal(grid, g[f"{name}_grid"])
    np.testing.assert_allclose(s, g[f"{name}_exp_svf"], atol=2e-5, rtol=1e-5)
    assert np.allclose(no.trapezoid_fov_mask(2 * H, W)[:H, :W], g[f"{name}_fov"])


def test_splat_oracle_matches_golden(golden):
    g = golden("splat.npz")
    depth, p2p, feats = synth.splat_inputs()
    xyz, xy, mask = co.frustum_to_bev(depth, p2p, [-12.8, -12.8, -2, 12.8, 12.8, 1], [0.1, 0.1])
    assert np.array_equal(xyz.view(np.uint32), g["xyz"].view(np.uint32))
    assert np.array_equal(xy.view(np.uint32), g["xy"].view(np.uint32))
    assert np.array_equal(mask, g["mask"])
    vol, dens, idx, _ = co.splat_soft(xy, feats * mask[:, None, :], 256, 256)
    XY = g["XY"]
    inb = (XY[..., 0] >= 0) & (XY[..., 0] < 256) & (XY[..., 1] >= 0) & (XY[..., 1] < 256)
    assert np.array_equal(idx[..., 0][inb], (XY[..., 1] * 256 + XY[..., 0])[inb])
    np.testing.assert_allclose(dens, g["dens"], atol=1e-6)
    np.testing.assert_allclose(vol.transpose(0, 2, 1)[g["nz"]], g["vol_nz"], atol=1e-5, rtol=1e-5)


def test_lidar_depth_loss_oracles_match_golden(golden):
    g = golden("lidar.npz")
    dm, dmm = co.lidar_raster(synth.os1_scan(seed=3)[::8], synth.lidar2camrect(128, 240), 128, 240)
    assert np.array_equal(dm, g["depth_m"]) and np.array_equal(dmm, g["depth_mm"].astype(np.float32))
    g = golden("depth.npz")
    m, b = co.depth_expectation(synth.depth_logits_inputs())
    assert np.array_equal(b, g["bins"])
    np.testing.assert_allclose(m, g["metric"], atol=1e-4)
    g = golden("loss.npz")
    expert, cfs, exp_svf, reward = synth.loss_inputs()
    cnt = co.expert_visitation(expert[:, :, :2, 2], 2, 64, 128, False)
    assert np.array_equal(cnt, g["counts"])
    fov = np.broadcast_to(g["fov"], (4,) + g["fov"].shape)
    loss, a, b_ = no.maxent_irl_loss_value(exp_svf, expert[:, :, :2, 2], fov, reward[:, 0], cfs)
    assert abs(a - float(g["mean_exp"])) < 1e-5 and abs(b_ - float(g["mean_svf"])) < 1e-5
    # reward_penalty term is zero in the golden (reward does not require grad there)
    assert abs(loss - float(g["loss"])) < 1e-5


@pytest.mark.parametrize("prof", ["peaky", "soft"])
def test_net_oracle_matches_golden_forward(golden, prof):
    """Well-conditioned tensors are compared tightly; the costmap only to the conditioning
    yardstick (the golden was produced on the build container's CPU)."""
    import creste_public_b200 as cb
    g = golden(f"forward_{prof}_64x96.npz")
    m = cb.build_maxentirl(image_size=(64, 96)).eval()
    assert len(m.state_dict()) == int(g["n_keys"])
    sd = synth.seeded_state_dict(m.state_dict(), 0, prof)
    rgbd, p2p = synth.net_inputs(64, 96, 1)
    out = no.forward(sd, rgbd, p2p)
    assert np.array_equal(out["depth_preds_bins"].numpy(), g["depth_bins"].astype(np.int64))
    f = out["depth_preds_feats"][0, ::16].numpy()
    assert np.abs(f - g["feats_sample"]).max() <= 1e-5 * np.abs(g["feats_sample"]).max()
    ref64 = no.forward(sd, rgbd, p2p, encoder_fp64=True)
    yard = float((ref64["traversability_preds"] - out["traversability_preds"]).abs().max())
    assert np.abs(out["traversability_preds"].numpy() - g["costmap"]).max() <= 10 * yard + 1e-4


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_oracle_pinned_against_live_reference():
    """Bit-level pin of the C oracle against the reference's own functions (build container)."""
    from oracle import ref_harness as rh
    vin = rh.build_ref_vin()
    r = torch.from_numpy(synth.vi_inputs(77, 2, 20, 28))
    v, pol, q = vin.value_iteration_manual(r, None, 0.001, 0.99)
    v0, q0, pi0, K = co.vi_solve(r.numpy())
    assert np.array_equal(v.numpy()[:, 0].view(np.uint32), v0.view(np.uint32))
    assert np.array_equal(q.numpy().view(np.uint32), q0.view(np.uint32))
    mods = rh.ref_modules()
    pc = synth.os1_scan(5)[::16]
    P = synth.lidar2camrect(64, 96)
    pts, dep = mods["projection"].pixels_to_depth(pc, {"lidar2camrect": P}, 64, 96)
    img = np.zeros((64, 96), np.float32)
    img[pts[:, 1], pts[:, 0]] = dep
    assert np.array_equal(img, co.lidar_raster(pc, P, 64, 96)[0])


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_mirror_state_dict_equals_reference():
    """Drop-in surface: identical parameter / buffer names, order, shapes and analytic buffers."""
    import creste_public_b200 as cb
    from oracle import ref_harness as rh
    ref, _ = rh.build_ref_maxentirl(image_size=(64, 96))
    ours = cb.build_maxentirl(image_size=(64, 96))
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
    for k in ("traversability_head.w", "transition_probs", "dynamics", "backbone.cam2map.lidar2map",
              "backbone.cam2map.grid_size", "backbone.cam2map.voxel_size"):
        assert torch.equal(a[k], b[k]), k
    assert torch.equal(ref.fov_mask, ours.fov_mask)
    ours.load_state_dict(a)   # a reference checkpoint loads strictly


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present (GPU box)")
def test_net_oracle_equals_live_reference_forward():
    from oracle import ref_harness as rh
    ref, _ = rh.build_ref_maxentirl(image_size=(64, 96))
    ref.eval()
    sd = synth.seeded_state_dict(ref.state_dict(), 3, "peaky")
    ref.load_state_dict(sd)
    rgbd, p2p = synth.net_inputs(64, 96, 1, seed=3)
    with torch.no_grad():
        a = ref((rgbd, p2p))
    b = no.forward(sd, rgbd, p2p)
    for k in ("depth_preds_logits", "depth_preds_feats", "depth_preds_metric", "dino_pe_feats"):
        assert torch.equal(a[k], b[k].view_as(a[k])), k
    assert torch.equal(a["depth_preds_bins"], b["depth_preds_bins"])
    assert float((a["traversability_preds"] - b["traversability_preds"]).abs().max()) < 1e-4


def test_splat_backward_restatement_matches_reference_golden(golden):
    """Stage-2 groundwork: the numpy restatement of splat_soft's backward -- the formulas a
    future creste_splat_soft_bwd kernel implements -- against the reference's own autograd (golden minted by
    oracle/gen_golden.py from splat_projection.py:262-354 on the seeded splat case)."""
    from oracle import splat_bwd_oracle as sb
    g = golden("splat_bwd.npz")
    H, W = int(g["grid"][0]), int(g["grid"][1])
    N, Cc, _ = g["feats"].shape
    rng = np.random.default_rng(int(g["g_seed"]))
    G = rng.standard_normal((N, Cc, H * W)).astype(np.float32)
    Gd = rng.standard_normal((N, H * W, 1)).astype(np.float32)[:, :, 0]
    dfe, dxy = sb.splat_backward(g["xy"], g["feats"], G, Gd, H, W)
    assert np.abs(dfe - g["dfeats"]).max() <= 1e-5 * np.abs(g["dfeats"]).max()
    assert np.abs(dxy - g["dxy"]).max() <= 1e-5 * np.abs(g["dxy"]).max()
    # and the forward half of the restatement against the forward golden of the same case
    f = golden("splat.npz")
    out, dens, _ = sb.splat_forward(g["xy"], g["feats"], H, W)
    np.testing.assert_allclose(dens, f["dens"], rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_bev_port_matches_reference():
    """Stage-2 groundwork: the train-mode BEV decoder port (torchvision ResNet-18 layers + DeconvHeads) equals the
    unmodified reference bit for bit on loss, all 78 gradients and the BatchNorm running statistics."""
    from oracle import bev_oracle as bo
    case = bo.make_case()
    ref, port = bo.reference_step(case), bo.port_step(case)
    assert ref["loss"] == port["loss"] and set(ref["grads"]) == set(port["grads"]) and len(ref["grads"]) == 78
    for k in ref["grads"]:
        assert np.array_equal(ref["grads"][k], port["grads"][k]), k
    for k in ref["buffers"]:
        assert np.array_equal(ref["buffers"][k], port["buffers"][k]), k


def test_bev_port_matches_golden(golden):
    from oracle import bev_oracle as bo
    g = golden("bev_step.npz")
    port = bo.port_step(bo.make_case())
    np.testing.assert_allclose(port["loss"], g["loss"], rtol=1e-5)
    names = [str(n) for n in g["grad_names"]]
    l2 = np.array([np.sqrt((port["grads"][n].astype(np.float64) ** 2).sum()) for n in names])
    np.testing.assert_allclose(l2, g["grad_l2"], rtol=1e-4)
    for n in ("layer2.0.conv1.weight", "layer2.0.downsample.0.weight"):
        ref = g["grad::" + n]
        assert np.abs(port["grads"][n] - ref).max() <= 1e-4 * np.abs(ref).max(), n
    np.testing.assert_allclose(port["buffers"]["bn1.running_mean"], g["bn1_running_mean"], rtol=1e-5, atol=1e-7)
