"""GPU parity tests of the HBM/latency-bound kernels: CUDA (through the C ABI) vs the C oracle
on the same seeded inputs, and vs the golden fixtures minted from the unmodified reference.

Tolerances: integer outputs exact (sweep count K, voxel indices, rollout
states, raster pixels, depth bins); v / q bit-exact (same fp32 operation order); pi, exp_svf,
bev features <= 1e-5 (+ relative 1e-5 where values exceed 1).
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import synth

pytestmark = pytest.mark.gpu


def _ops():
    from creste_public_b200 import ops
    return ops


# ------------------------------------------------------------------------------------------ VI
@pytest.mark.parametrize("name", ["b2_16x16", "b1_64x64", "b3_24x40"])
def test_vi_matches_reference_golden(cuda, golden, name):
    g = golden("vi.npz")
    seed, B, H, W, K = [int(x) for x in g[f"{name}_meta"]]
    r = torch.from_numpy(synth.vi_inputs(seed, B, H, W)).to(cuda)
    v, q, pi, info = _ops().vi_solve(r, 0.99, 1e-3)
    assert int(info[0]) == K and int(info[1]) == 0
    assert np.array_equal(v.cpu().numpy().view(np.uint32), g[f"{name}_v"].view(np.uint32))
    if f"{name}_q" in g.files:
        assert np.array_equal(q.cpu().numpy().view(np.uint32), g[f"{name}_q"].view(np.uint32))
        np.testing.assert_allclose(pi.cpu().numpy(), g[f"{name}_pi"], atol=1e-6, rtol=0)


@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (8, 64, 128), (1, 5, 7), (3, 33, 17),
                                   (2, 128, 128), (8, 256, 256)])
def test_vi_matches_oracle(cuda, B, H, W):
    r = synth.vi_inputs(100 + B + H, B, H, W)
    v0, q0, pi0, K0 = co.vi_solve(r)
    v, q, pi, info = _ops().vi_solve(torch.from_numpy(r).to(cuda), 0.99, 1e-3)
    assert int(info[0]) == K0
    assert np.array_equal(v.cpu().numpy()[:, 0].view(np.uint32), v0.view(np.uint32))
    assert np.array_equal(q.cpu().numpy().view(np.uint32), q0.view(np.uint32))
    np.testing.assert_allclose(pi.cpu().numpy(), pi0, atol=1e-6, rtol=0)


@pytest.mark.parametrize("B,H,W,gamma,thr,ms", [
    (4, 256, 256, 0.99, 1e-3, 4096),      # cluster of 16 per sample
    (2, 100, 64, 0.99, 1e-3, 4096),       # ragged last strip
    (16, 64, 128, 0.99, 1e-3, 4096),      # many small clusters
    (5, 48, 256, 0.95, 1e-3, 4096),
    (1, 256, 256, 0.99, 1e-3, 4096),
    (3, 32, 32, 0.5, 1e-1, 4096),         # K smaller than the snapshot period
    (3, 32, 32, 0.9, 1e-2, 4096),         # K between snapshots
    (2, 64, 64, 0.99, 1e-3, 37),          # max_sweeps hit
    (2, 64, 64, 0.99, 1e-3, 1),
    (64, 64, 64, 0.99, 1e-3, 4096),       # more samples than co-resident clusters
])
def test_vi_strip_kernel_cases(cuda, B, H, W, gamma, thr, ms):
    """Checkpoint + replay and the lagged stop decision must reproduce the oracle's sweep count
    and value function bit for bit, whatever K is relative to the snapshot period."""
    r = synth.vi_inputs(7 * B + H + W, B, H, W)
    v0, q0, pi0, K0 = co.vi_solve(r, gamma, thr, ms)
    v, q, pi, info = _ops().vi_solve(torch.from_numpy(r).to(cuda), gamma, thr, ms)
    assert int(info[0]) == K0, (int(info[0]), K0)
    assert np.array_equal(v.cpu().numpy()[:, 0].view(np.uint32), v0.view(np.uint32))
    assert np.array_equal(q.cpu().numpy().view(np.uint32), q0.view(np.uint32))
    np.testing.assert_allclose(pi.cpu().numpy(), pi0, atol=1e-6, rtol=0)


def test_vi_constant_reward_known_answer(cuda):
    # interior cells of a constant-reward grid: v_k = c * sum_{i<k} gamma^i (all 8 actions equal)
    r = torch.full((1, 1, 40, 40), 0.5, device=cuda)
    v, q, pi, info = _ops().vi_solve(r, 0.9, 1e-4)
    K = int(info[0])
    expect = 0.5 * (1 - 0.9 ** K) / (1 - 0.9)
    assert abs(float(v[0, 0, 20, 20]) - expect) < 1e-3
    np.testing.assert_allclose(pi[0, :, 20, 20].cpu().numpy(), np.full(8, 0.125), atol=1e-6)


def test_vi_max_sweeps_flag(cuda):
    r = torch.rand(1, 1, 16, 16, device=cuda)
    v, q, pi, info = _ops().vi_solve(r, 0.99, 1e-3, max_sweeps=5)
    assert int(info[0]) == 5 and int(info[1]) == 1


# ----------------------------------------------------------------------------------------- SVF
@pytest.mark.parametrize("name", ["b2_32x64", "b2_32x64_zt", "b1_64x128"])
def test_svf_matches_reference_golden(cuda, golden, name):
    g = golden("svf.npz")
    seed, B, H, W, T, zt = [int(x) for x in g[f"{name}_meta"]]
    r, expert = synth.svf_inputs(seed, B, H, W, T)
    ops = _ops()
    v, q, pi, info = ops.vi_solve(torch.from_numpy(r).to(cuda), 0.99, 1e-3)
    rc = torch.from_numpy(expert[:, :, :2, 2].copy()).to(cuda)
    fov = torch.from_numpy(g[f"{name}_fov"]).to(cuda)
    svf, states, grid = ops.svf(pi, rc, fov, T, 2, True, 0.005, bool(zt))
    assert np.array_equal(states.cpu().numpy(), g[f"{name}_states"])
    assert np.array_equal(grid.cpu().numpy(), g[f"{name}_grid"])
    np.testing.assert_allclose(svf.cpu().numpy(), g[f"{name}_exp_svf"], atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("B,H,W,T", [(4, 64, 128, 50), (2, 256, 256, 50), (3, 20, 24, 12)])
def test_svf_matches_oracle_and_conserves_mass(cuda, B, H, W, T):
    from oracle import net_oracle
    r, expert = synth.svf_inputs(7, B, H, W, T)
    ops = _ops()
    v, q, pi, info = ops.vi_solve(torch.from_numpy(r).to(cuda), 0.99, 1e-3)
    fov = net_oracle.trapezoid_fov_mask(2 * H, W)[:H, :W]
    rc = expert[:, :, :2, 2].copy()
    for sharpen in (True, False):
        s0, st0, g0 = co.svf(pi.cpu().numpy(), rc, fov, T, 2, sharpen, 0.005, False)
        s, st, g = ops.svf(pi, torch.from_numpy(rc).to(cuda), torch.from_numpy(fov).to(cuda), T, 2,
                           sharpen, 0.005, False)
        assert np.array_equal(st.cpu().numpy(), st0)
        assert np.array_equal(g.cpu().numpy(), g0)
        np.testing.assert_allclose(s.cpu().numpy(), s0, atol=2e-5, rtol=1e-5)
    assert float(s.sum()) <= B * T + 1e-2


# --------------------------------------------------------------------------------------- splat
def test_frustum_and_splat_match_reference_golden(cuda, golden):
    g = golden("splat.npz")
    depth, p2p, feats = synth.splat_inputs()
    N, Hs, Ws = depth.shape
    ops = _ops()
    rng = [-12.8, -12.8, -2.0, 12.8, 12.8, 1.0]
    xy, z, mask = ops.frustum_to_bev(torch.from_numpy(depth).to(cuda), torch.from_numpy(p2p).to(cuda),
                                     rng, [0.1, 0.1])
    assert np.array_equal(xy.cpu().numpy().view(np.uint32), g["xy"].view(np.uint32))
    assert np.array_equal(z.cpu().numpy().view(np.uint32), g["xyz"][:, 2].view(np.uint32))
    assert np.array_equal(mask.cpu().numpy().astype(bool), g["mask"])
    f = torch.from_numpy(feats).to(cuda).permute(0, 2, 1).contiguous()  # [N,P,F]
    out = ops.splat_soft(xy, f, mask, 256, 256, want_idx=True)
    XY = g["XY"]
    idx = out["idx"].cpu().numpy()
    # bit-exact voxel indices: tap 0 of every in-bounds point is Y0*W + X0
    inb = (XY[..., 0] >= 0) & (XY[..., 0] < 256) & (XY[..., 1] >= 0) & (XY[..., 1] < 256)
    assert np.array_equal(idx[..., 0][inb], (XY[..., 1] * 256 + XY[..., 0])[inb])
    assert np.all(idx[..., 0][~inb] == -1)
    dens = out["dens"].cpu().numpy().reshape(N, -1)
    np.testing.assert_allclose(dens, g["dens"], atol=1e-5, rtol=1e-5)
    vol = out["bev_nchw"].cpu().numpy().reshape(N, feats.shape[1], -1).transpose(0, 2, 1)
    np.testing.assert_allclose(vol[g["nz"]], g["vol_nz"], atol=1e-5, rtol=1e-5)
    assert np.all(vol[~g["nz"]] == 0)
    nhwc = out["bev_nhwc"].cpu().numpy().reshape(N, -1, feats.shape[1])
    assert np.array_equal(nhwc, vol)


@pytest.mark.parametrize("N,Hs,Ws,F", [(1, 128, 240, 96), (2, 32, 60, 8)])
def test_splat_matches_oracle(cuda, N, Hs, Ws, F):
    g = np.random.default_rng(9)
    depth = (g.random((N, Hs, Ws), dtype=np.float32) * 25 + 0.3).astype(np.float32)
    p2p = np.stack([synth.make_p2p(Hs * 4, Ws * 4)] * N)
    feats = g.standard_normal((N, Hs * Ws, F)).astype(np.float32)
    rng = [-12.8, -12.8, -2.0, 12.8, 12.8, 1.0]
    xyz0, xy0, m0 = co.frustum_to_bev(depth, p2p, rng, [0.1, 0.1])
    fm = feats * m0[..., None]
    vol0, dens0, idx0, _ = co.splat_soft(xy0, fm.transpose(0, 2, 1), 256, 256)
    ops = _ops()
    xy, z, mask = ops.frustum_to_bev(torch.from_numpy(depth).to(cuda), torch.from_numpy(p2p).to(cuda),
                                     rng, [0.1, 0.1])
    assert np.array_equal(xy.cpu().numpy().view(np.uint32), xy0.view(np.uint32))
    assert np.array_equal(mask.cpu().numpy().astype(bool), m0)
    out = ops.splat_soft(xy, torch.from_numpy(feats).to(cuda), mask, 256, 256, want_idx=True)
    assert np.array_equal(out["idx"].cpu().numpy(), idx0)          # bit-exact voxel indices
    np.testing.assert_allclose(out["dens"].cpu().numpy().reshape(N, -1), dens0, atol=1e-5, rtol=1e-5)
    np.testing.assert_allclose(out["bev_nchw"].cpu().numpy().reshape(N, F, -1), vol0, atol=1e-5,
                               rtol=1e-5)


def test_splat_one_hot_point(cuda):
    # a point exactly at a cell corner deposits weight {1,0,0,0} (analytic known answer)
    xy = torch.tensor([[[10.0, 20.0], [300.0, 5.0], [-0.5, 0.25]]], device=cuda)
    f = torch.ones(1, 3, 4, device=cuda)
    out = _ops().splat_soft(xy, f, None, 256, 256, want_idx=True)
    d = out["dens"][0, 0]
    assert float(d[20, 10]) == 1.0
    idx = out["idx"][0].cpu().numpy()
    assert list(idx[0]) == [20 * 256 + 10, 21 * 256 + 10, 20 * 256 + 11, 21 * 256 + 11]
    assert list(idx[1]) == [-1, -1, -1, -1]
    assert list(idx[2]) == [-1, -1, 0, 256]
    assert abs(float(d[0, 0]) - 0.5 * 0.75) < 1e-7 and abs(float(d[1, 0]) - 0.5 * 0.25) < 1e-7


# ---------------------------------------------------------------------------- LiDAR, depth, loss
def test_lidar_raster_matches_reference_golden(cuda, golden):
    g = golden("lidar.npz")
    H, W = 128, 240
    pc = synth.os1_scan(seed=3)[::8]
    dm, dmm = _ops().lidar_raster(torch.from_numpy(pc).to(cuda), synth.lidar2camrect(H, W), H, W)
    assert np.array_equal(dm.cpu().numpy(), g["depth_m"])
    assert np.array_equal(dmm.cpu().numpy(), g["depth_mm"].astype(np.float32))


def test_lidar_raster_full_size_matches_oracle(cuda):
    H, W = 512, 960
    pc = synth.os1_scan(seed=0)
    P = synth.lidar2camrect(H, W)
    dm0, dmm0 = co.lidar_raster(pc, P, H, W)
    dm, dmm = _ops().lidar_raster(torch.from_numpy(pc).to(cuda), P, H, W)
    assert np.array_equal(dm.cpu().numpy(), dm0) and np.array_equal(dmm.cpu().numpy(), dmm0)
    assert 0.02 < (dm0 > 0).mean() < 0.08
    # empty cloud -> empty raster
    dm, _ = _ops().lidar_raster(torch.zeros(0, 3, device=cuda), P, H, W)
    assert float(dm.abs().max()) == 0.0


def test_depth_expectation_matches_reference_golden(cuda, golden):
    g = golden("depth.npz")
    logits = synth.depth_logits_inputs()
    x = torch.from_numpy(logits).to(cuda).permute(0, 2, 3, 1).contiguous()
    m, b = _ops().depth_expectation(x)
    assert np.array_equal(b.cpu().numpy(), g["bins"])
    np.testing.assert_allclose(m.cpu().numpy(), g["metric"], atol=1e-4, rtol=0)
    # one-hot row -> exactly its linspace value
    oh = torch.zeros(1, 1, 2, 128, device=cuda)
    oh[0, 0, 0, 5] = 200.0
    oh[0, 0, 1, 127] = 200.0
    m, b = _ops().depth_expectation(oh)
    vals = torch.linspace(300, 25600, 128)
    assert abs(float(m[0, 0, 0]) - float(vals[5]) / 1000) < 1e-6
    assert abs(float(m[0, 0, 1]) - 25.6) < 1e-6 and int(b[0, 0, 1]) == 127


def test_expert_visitation_matches_reference_golden(cuda, golden):
    g = golden("loss.npz")
    expert, cfs, exp_svf, reward = synth.loss_inputs()
    rc = expert[:, :, :2, 2].copy()
    ms = int(np.ceil(np.linalg.norm((rc[:, 1:] - rc[:, :-1]) / np.float32(2), axis=-1)).max())
    cnt = _ops().expert_visitation(torch.from_numpy(rc).to(cuda), 2, ms, 64, 128)
    assert np.array_equal(cnt.cpu().numpy(), g["counts"])
    for cf in cfs:
        if cf is None:
            continue
        t = cf["trajectories"]
        ms = int(np.ceil(np.linalg.norm((t[:, 1:] - t[:, :-1]) / 2.0, axis=-1)).max())
        c = _ops().expert_visitation(torch.from_numpy(t).to(cuda), 2, ms, 64, 128)
        assert np.array_equal(c.cpu().numpy(), co.expert_visitation(t, 2, 64, 128, True))


def test_projection_mirror_pixels_to_depth(cuda, golden):
    """creste.utils.projection.pixels_to_depth mirror (reference signature) vs the golden raster."""
    from creste_public_b200.creste.utils import projection as pj
    g = golden("lidar.npz")
    H, W = 128, 240
    pc = synth.os1_scan(seed=3)[::8]
    pts, dep = pj.pixels_to_depth(pc, {"lidar2camrect": synth.lidar2camrect(H, W)}, H, W)
    img = np.zeros((H, W), np.float32)
    img[pts[:, 1], pts[:, 0]] = dep
    assert np.array_equal(img, g["depth_m"])
    rgbd = pj.make_rgbd(torch.zeros(3, H, W), pc, synth.lidar2camrect(H, W))
    assert np.array_equal(rgbd[3].cpu().numpy(), g["depth_mm"].astype(np.float32))


def test_os1_file_to_depth_png_pipeline(cuda, tmp_path):
    """Formats either side of the raster kernel: an OS1 `.bin` sweep (5 float32 per point) -> creste_lidar_raster ->
    the uint16-mm depth PNG the reference's preprocessing writes -> read back as the network's 4th channel: identical
    to the C oracle of projection.py:64-134 + build_dense_depth.py:461-463."""
    from creste_public_b200.creste.datasets import coda_formats as fmt
    H, W = 512, 960
    pc5 = np.concatenate([synth.os1_scan(4), np.ones((131072, 2), np.float32)], axis=1)
    fmt.write_os1_bin(str(tmp_path / "sweep.bin"), pc5)
    pc = torch.from_numpy(fmt.read_os1_bin(str(tmp_path / "sweep.bin"))).cuda()
    assert tuple(pc.shape) == (131072, 5)
    _, dmm = _ops().lidar_raster(pc, synth.lidar2camrect(H, W), H, W, want_m=False)
    fmt.write_depth_png(str(tmp_path / "0.png"), dmm.cpu().numpy())
    back = fmt.read_depth_png(str(tmp_path / "0.png"))
    _, want = co.lidar_raster(pc5, synth.lidar2camrect(H, W), H, W)
    assert np.array_equal(back, want)
