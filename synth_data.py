"""Seeded synthetic parameters and inputs -- pure generators, no
algorithm of the path: shared by bench.py (both arms), the tests and the oracle.

No dataset or checkpoint can be downloaded, so every workload uses synthetic inputs of the
reference's shapes and seeded random weights of the reference's architecture.  Parameters are
drawn per state-dict key from a numpy Generator seeded by crc32(key) -- independent of torch's
module-construction order and identical in this container and on the GPU box -- so the reference
model, the oracle and the CUDA product can all be given the very same weights by name.
"""
import zlib

import numpy as np
import torch

FIXED_BUFFERS = (
    "dynamics", "transition_probs", "traversability_head.w",
    "point_cloud_range", "max_bound", "min_bound", "voxel_size", "grid_size", "lidar2map",
)


def _rng(key, seed):
    return np.random.default_rng((zlib.crc32(key.encode()) + 7919 * seed) & 0xFFFFFFFF)


def seeded_state_dict(template, seed=0, depth_profile="peaky"):
    """template: mapping name -> tensor (only shapes/dtypes are used).  Returns a new dict.

    depth_profile: "peaky" -- depth-head BN gain U[3,6]: near one-hot bin distributions as a
    trained depth classifier produces (argmax-like depth, ill-conditioned w.r.t. the logits);
    "soft" -- gain U[0.5,1.5] with a bias tilt toward the near bins: smooth expectation, used
    by the end-to-end 1e-4 parity test (see DESIGN.md "Conditioning of the end-to-end check").
    """
    out = {}
    for k, t in template.items():
        if any(k == f or k.endswith("." + f) or k.endswith(f) for f in FIXED_BUFFERS):
            out[k] = t.clone()
            continue
        g = _rng(k, seed)
        shape = tuple(t.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(t)
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy(g.normal(0, 0.1, shape).astype(np.float32))
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(g.uniform(0.5, 1.5, shape).astype(np.float32))
        elif t.ndim == 1 and (k[: k.rfind(".")] + ".running_mean") in template:
            # depth head: peaky bins so the synthetic depth spans the range; reward head:
            # costmap in roughly [0, 1] with most cells positive (a trained reward map's range)
            if k.endswith(".weight"):
                lo, hi = (0.5, 1.5)
                if "depth_head" in k and depth_profile == "peaky":
                    lo, hi = 3.0, 6.0
                elif "r.postpool.0.norm" in k:
                    lo, hi = 0.05, 0.15
                out[k] = torch.from_numpy(g.uniform(lo, hi, shape).astype(np.float32))
            else:
                mean = 0.3 if "r.postpool.0.norm" in k else 0.0
                b = g.normal(mean, 0.1, shape).astype(np.float32)
                if "depth_head" in k and depth_profile == "soft":
                    kk = np.arange(shape[0], dtype=np.float32)
                    b += np.where(kk < 48, 5.0 * (1.0 - kk / 48.0), -1.0).astype(np.float32)
                out[k] = torch.from_numpy(b)
        elif t.ndim >= 2:
            fan_in = int(np.prod(shape[1:]))
            std = np.sqrt(2.0 / fan_in)
            w = (g.standard_normal(shape) * std).astype(np.float32)
            if k.endswith("_conv_stem.weight") and shape[1] == 4:
                # channel 3 is sparse LiDAR depth in un-normalised millimetres
                # (codapefree_dataloader.py:864-873): keep the synthetic activations O(1)
                w[:, 3] *= np.float32(2e-4)
            out[k] = torch.from_numpy(w)
        elif t.ndim == 1:
            out[k] = torch.from_numpy(g.normal(0, 0.05, shape).astype(np.float32))
        else:
            out[k] = t.clone()
    return out


# LiDAR (x fwd, y left, z up) <- camera (x right, y down, z fwd), small lever arm.
T_CAM_TO_LIDAR = np.array([[0, 0, 1, 0.10], [-1, 0, 0, 0.05], [0, -1, 0, -0.30], [0, 0, 0, 1.0]])


def intrinsics(H, W):
    """fx=fy=720, cx=480, cy=256 at 512x960, scaled with the image."""
    s = W / 960.0
    return np.array([[720.0 * s, 0, W / 2.0], [0, 720.0 * s, H / 2.0], [0, 0, 1.0]])


def make_p2p(H, W, ds=4):
    """creste/utils/projection.py:11-34 + codapefree_dataloader.py:803-816: pixel@1/ds-res ->
    LiDAR, built in float64 then cast to float32.  Returns [4,4] float32."""
    K = intrinsics(H, W)
    K[:2, :] /= ds
    P = np.eye(4)
    P[:3, :3] = np.linalg.inv(K)
    return (T_CAM_TO_LIDAR @ P).astype(np.float32)


def lidar2camrect(H, W):
    """[3,4] float64 LiDAR -> rectified-camera pixel projection for the rasteriser."""
    return intrinsics(H, W) @ np.linalg.inv(T_CAM_TO_LIDAR)[:3, :]


def os1_scan(seed=0, beams=128, azimuths=1024):
    """Synthetic Ouster OS1-128 sweep: 128 beams (+22.5..-22.5 deg) x 1024 azimuths,
    range U[1,25] m -> xyz float32 [131072, 3]."""
    g = np.random.default_rng(1000 + seed)
    el = np.deg2rad(np.linspace(22.5, -22.5, beams))
    az = np.linspace(-np.pi, np.pi, azimuths, endpoint=False)
    E, A = np.meshgrid(el, az, indexing="ij")
    R = g.uniform(1.0, 25.0, E.shape)
    pc = np.stack([R * np.cos(E) * np.cos(A), R * np.cos(E) * np.sin(A), R * np.sin(E)], -1)
    return pc.reshape(-1, 3).astype(np.float32)


def rgb_frames(B, H, W, seed=0):
    g = np.random.default_rng(2000 + seed)
    return g.random((B, 1, 3, H, W), dtype=np.float32)


def expert_poses(B, T=50, H_un=256, W_un=256, seed=0):
    """[B,T,3,3] SE(2) poses in un-pooled BEV cell units, heading 'north' from the bottom centre
    of the top half of the map (the part the reward net sees)."""
    g = np.random.default_rng(3000 + seed)
    e = np.zeros((B, T, 3, 3), np.float32)
    e[:, :, 0, 0] = e[:, :, 1, 1] = e[:, :, 2, 2] = 1
    for b in range(B):
        e[b, :, 0, 2] = np.linspace(H_un / 2 - 3, H_un * 0.12, T)
        e[b, :, 1, 2] = np.linspace(W_un / 2, W_un / 2 + g.uniform(-0.3, 0.3) * W_un, T)
    return e


def counterfactuals(expert, every=2, shift=30.0):
    """Per-sample list: dict(trajectories f64 [3,T,2], rank [0,1,1]) for every `every`-th sample,
    None otherwise (scripts/traversability/rlhf/app.py:201-224 pickle layout)."""
    out = []
    for b in range(expert.shape[0]):
        if b % every:
            out.append(None)
            continue
        rc = expert[b, :, :2, 2].astype(np.float64)
        out.append({"trajectories": np.stack([rc, rc + [0, shift], rc - [0, shift]]),
                    "rank": np.array([0, 1, 1])})
    return out


# ---- seeded inputs shared by oracle/gen_golden.py and tests/
def vi_inputs(seed, B, H, W):
    g = np.random.default_rng(seed)
    return g.random((B, 1, H, W), dtype=np.float32)


def splat_inputs(seed=5, N=2, Hs=16, Ws=24, F=8):
    g = np.random.default_rng(seed)
    depth = (g.random((N, Hs, Ws), dtype=np.float32) * 24 + 0.3).astype(np.float32)
    p2p = np.stack([make_p2p(Hs * 4, Ws * 4)] * N)
    p2p[1, :3, 3] += np.array([0.3, -0.2, 0.1], np.float32)
    feats = g.standard_normal((N, F, Hs * Ws)).astype(np.float32)
    return depth, p2p, feats


def svf_inputs(seed, B, H, W, T):
    r = vi_inputs(seed, B, H, W)
    expert = expert_poses(B, T, 2 * 2 * H, 2 * W, seed)  # un-pooled cells; top half = 2H rows
    return r, expert


def loss_inputs(B=4, Hm=64, Wm=128, T=50):
    """Inputs of the MaxEntIRLLoss golden case (exp_svf, reward, expert, counterfactuals)."""
    expert = expert_poses(B, T, 256, 256, seed=41)
    cfs = counterfactuals(expert)
    g = np.random.default_rng(42)
    exp_svf = g.random((B, Hm, Wm), dtype=np.float32)
    reward = g.random((B, 1, Hm, Wm), dtype=np.float32)
    return expert, cfs, exp_svf, reward


def depth_logits_inputs():
    g = np.random.default_rng(31)
    return np.maximum(g.standard_normal((2, 128, 6, 10)) * 3, 0).astype(np.float32)


def head_inputs(B, Hm, Wm, seed=0, T=50):
    """Synthetic BEV head predictions [B,{32,6,2},4Hm,2Wm] + expert / counterfactuals / FOV mask
    for an Hm x Wm reward grid (un-pooled BEV 4Hm x 2Wm)."""
    g = np.random.default_rng(7000 + seed)
    feat = [torch.from_numpy(g.standard_normal((B, c, 4 * Hm, 2 * Wm), dtype=np.float32))
            for c in (32, 6, 2)]
    expert = torch.from_numpy(expert_poses(B, T, 4 * Hm, 2 * Wm, seed))
    cfs = counterfactuals(expert.numpy(), every=2, shift=0.12 * 2 * Wm)
    fov = trapezoid_fov_mask(4 * Hm, 2 * Wm, 70, 70, 7 * Wm / 128.0, 200 * Wm / 128.0)
    fov = torch.from_numpy(np.ascontiguousarray(fov)).unsqueeze(0).repeat(B, 1, 1)
    return feat, expert, fov, cfs


def stage1_loss_inputs(seed=51, B=2, H=16, W=24, D=128, Z=8):
    """Inputs of the stage-1 validation losses: depth logits, their arg-max bins, a depth label in
    millimetres (20 % invalid zeros, some beyond the range), DINO predictions / targets with
    non-finite targets (pixels without a feature)."""
    g = np.random.default_rng(seed)
    logits = np.maximum(g.standard_normal((B, D, H, W)) * 3, 0).astype(np.float32)
    label = g.uniform(300, 25600, (B, 1, H, W)).astype(np.float32)
    label[g.random((B, 1, H, W)) < 0.2] = 0
    label[0, 0, 0, :4] = [25600.0, 25599.9, 300.0, 26000.0]
    # put some labels next to the arg-max bin so that the accuracy is not ~0
    am = logits.argmax(1)
    near = g.random((B, H, W)) < 0.3
    label[:, 0][near] = (300 + (am[near] + 0.5) * (25300.0 / D)).astype(np.float32)
    pred = g.standard_normal((B, 1, Z, H, W)).astype(np.float32)
    gt = g.standard_normal((B, 1, Z, H, W)).astype(np.float32)
    gt[g.random(gt.shape) < 0.1] = np.inf
    return logits, label, pred, gt


def trapezoid_fov_mask(H, W, top=70, bottom=70, near=0, far=100):
    """Synthetic field-of-view mask input: the trapezoid of creste/utils/train_utils.py:511-557 in
    numpy float32 (the product builds its own with creste.utils.train_utils)."""
    import math
    y, x = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    cx, cy = W / 2, H / 2
    dx = (x - cx).astype(np.float32)
    dy = (y - cy).astype(np.float32)
    dist = np.sqrt(dx ** 2 + dy ** 2).astype(np.float32)
    ang = (np.arctan2(dx, -dy).astype(np.float32) * np.float32(180) / np.float32(math.pi))
    t, b = np.float32(top / 2), np.float32(bottom / 2)
    spread = np.where(dist <= near, t, np.where(dist >= far, b,
                      t + (b - t) * ((dist - near) / np.float32(far - near))))
    return (dist >= near) & (dist <= far) & (np.abs(ang) <= spread)


def distill_batch(B, H, W, seed=0, Z=128):
    """One synthetic stage-1 batch: `image` [B,1,4,H,W] (RGB in [0,1) +
    ~4 % filled LiDAR depth in mm), `depth_label` [B,1,H/4,W/4] mm in [300, 25600) with 20 % invalid
    zeros, `fimg_label` [B,1,Z,H/4,W/4] ~ N(0,1) DINO targets."""
    g = torch.Generator().manual_seed(4000 + seed)
    rgb = torch.rand(B, 1, 3, H, W, generator=g)
    depth = torch.rand(B, 1, 1, H, W, generator=g) * 25300.0 + 300.0
    depth = depth * (torch.rand(B, 1, 1, H, W, generator=g) < 0.04)
    lab = torch.rand(B, 1, H // 4, W // 4, generator=g) * 25300.0 + 300.0
    lab = lab * (torch.rand(B, 1, H // 4, W // 4, generator=g) >= 0.2)
    fimg = torch.randn(B, 1, Z, H // 4, W // 4, generator=g)
    return {"image": torch.cat([rgb, depth], dim=2), "depth_label": lab, "fimg_label": fimg}


def ssc_batch(B, H, W, seed=0, G=256):
    """Synthetic stage-2 (train_ssc.py) batch, shapes of the reference's CODa loader: RGB-D image + p2p, the stage-1
    labels (sparse LiDAR depth at 1/4 resolution, DINO feature targets) and the BEV labels on the G x G grid -- SAM
    mask ids (blocky regions, 0 = unlabeled), dynamic-object classes in channel 1, (min, max) elevation with
    unobserved (NaN) cells, and a trapezoidal FOV mask."""
    import torch
    g = np.random.default_rng(4200 + seed)
    batch = distill_batch(B, H, W, seed)
    batch["p2p"] = torch.from_numpy(make_p2p(H, W)).view(1, 1, 4, 4).repeat(B, 1, 1, 1)
    # SAM masks: a few dozen segments per frame over a mostly unlabeled grid (the contrastive loss is O(N^2) in the
    # number of sampled labelled cells: ~15 % coverage x 20 ids keeps N at the 2-3e4 rows of a real CODa batch
    # rather than every cell of the map)
    cell = G // 16
    ids = g.integers(1, 21, (B, 16, 16)) * (g.random((B, 16, 16)) < 0.15)
    sam = np.repeat(np.repeat(ids, cell, axis=1), cell, axis=2)
    batch["3d_sam_label"] = torch.from_numpy(sam[:, None].astype(np.int64))
    dyn = np.zeros((B, 2, G, G), np.float32)
    dyn[:, 1] = np.repeat(np.repeat(g.integers(0, 6, (B, 32, 32)), G // 32, axis=1), G // 32, axis=2)
    batch["3d_sam_dynamic_label"] = torch.from_numpy(dyn)
    elev = g.standard_normal((B, 2, G, G)).astype(np.float32) * 0.3
    elev[:, 1] = elev[:, 0] + np.abs(elev[:, 1])
    elev[g.random((B, 2, G, G)) < 0.3] = np.nan
    batch["elevation_label"] = torch.from_numpy(elev)
    fov = trapezoid_fov_mask(G, G, 70, 70, 7, 200)
    batch["fov_mask"] = torch.from_numpy(np.broadcast_to(np.asarray(fov, bool), (B, G, G)).copy())
    return batch
