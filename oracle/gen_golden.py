"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under the
shim set.  TEST INFRASTRUCTURE; run in the build container only:

    python -m oracle.gen_golden

The fixtures pin (a) the oracle (oracle/creste_oracle.c, oracle/net_oracle.py) on CPU and
(b) the CUDA product on the GPU box, where the reference tree does not exist.  Inputs are
regenerated from seeds by oracle/synth.py at test time; only small outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def irl_step_golden():
    from oracle import irl_oracle
    case = irl_oracle.make_case(seed=3, B=2, H=8, W=16)
    ref = irl_oracle.reference_step(case, steps=2)
    d = {"loss": ref["loss"][0], "loss2": ref["loss"][1], "reward_penalty": ref["reward_penalty"][0],
         "mean_exp": ref["mean_expected_svf_rewards"][0], "mean_svf": ref["mean_svf_rewards"][0],
         "sum_cf": ref["sum_cf_rewards"][0], "sum_opt": ref["sum_opt_rewards"][0], "r": ref["r"]}
    d.update({"grad/" + k: v for k, v in ref["grads"].items()})
    d.update({"param2/" + k: v for k, v in ref["params"].items()})
    np.savez_compressed(os.path.join(OUT, "irl_step.npz"), **d)


def stage1_loss_golden():
    mods = rh.ref_modules()
    from omegaconf import OmegaConf
    cfgs = rh.compose_cfgs()
    lu = mods["loss_utils"]
    logits, label, pred, gt = [torch.from_numpy(a) for a in synth.stage1_loss_inputs()]
    td = {"outputs/depth_preds_logits": logits, "outputs/depth_preds_bins": logits.argmax(1),
          "inputs/depth_label": label, "outputs/dino_pe_feats": pred, "inputs/fimg_label": gt}
    out = {}
    for lc in cfgs["distill"]["loss"]:
        L = getattr(lu, lc["name"])(OmegaConf.create(lc))
        ld, md = L.loss(td)
        out.update({f"{lc['name']}/{k}": np.float32(v.item()) for k, v in ld.items()})
        out.update({f"{lc['name']}/{k}": np.float32(v.item()) for k, v in md.items()})
    np.savez_compressed(os.path.join(OUT, "stage1_losses.npz"), **out)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    mods = rh.ref_modules()
    from omegaconf import OmegaConf
    vin = rh.build_ref_vin()

    # ---- value iteration (vin.py:48-80)
    vi = {}
    for name, (seed, B, H, W) in {"b2_16x16": (11, 2, 16, 16), "b1_64x64": (0, 1, 64, 64),
                                  "b3_24x40": (12, 3, 24, 40)}.items():
        r = torch.from_numpy(synth.vi_inputs(seed, B, H, W))
        v, pol, q = vin.value_iteration_manual(r, None, threshold=0.001, discount=0.99)
        # count sweeps by replaying the loop condition with the reference's own conv
        vv = torch.zeros_like(r)
        K, delta = 0, float("inf")
        while delta > 0.001:
            qq = torch.nn.functional.conv2d(r + vv * 0.99, vin.w, stride=1, padding=1)
            nv = qq.max(dim=1, keepdim=True)[0]
            delta = (nv - vv).abs().max().item()
            vv = nv
            K += 1
        assert torch.equal(vv, v)
        vi[f"{name}_meta"] = np.array([seed, B, H, W, K])
        vi[f"{name}_v"] = v.numpy()
        if H <= 24:
            vi[f"{name}_q"] = q.numpy()
            vi[f"{name}_pi"] = pol.numpy()
    np.savez_compressed(os.path.join(OUT, "vi.npz"), **vi)

    # ---- SVF + rollout (lfd.py:156-277)
    svf = {}
    for name, (seed, B, H, W, T, zt) in {"b2_32x64": (21, 2, 32, 64, 20, False),
                                         "b2_32x64_zt": (21, 2, 32, 64, 20, True),
                                         "b1_64x128": (22, 1, 64, 128, 50, False)}.items():
        model, _ = rh.build_ref_maxentirl(image_size=(64, 96), solve_mdp=True, map_size=(H, W),
                                          action_horizon=T, zero_terminal_state=zt)
        r, expert = synth.svf_inputs(seed, B, H, W, T)
        v, pol, q = model.traversability_head.value_iteration_manual(
            torch.from_numpy(r), None, threshold=0.001, discount=0.99)
        out = model.expected_state_visitation_frequency(pol.clone(), torch.from_numpy(expert))
        svf[f"{name}_meta"] = np.array([seed, B, H, W, T, int(zt)])
        svf[f"{name}_fov"] = model.fov_mask[0, 0].numpy()
        svf[f"{name}_exp_svf"] = out["exp_svf"].numpy()
        svf[f"{name}_states"] = out["state_preds"].numpy()
        svf[f"{name}_grid"] = out["state_preds_grid"].numpy()
    np.savez_compressed(os.path.join(OUT, "svf.npz"), **svf)

    # ---- frustum -> BEV splat (splat_projection.py)
    model, cfgs = rh.build_ref_maxentirl(image_size=(64, 96))
    c2m = model.backbone.cam2map
    depth, p2p, feats = synth.splat_inputs()
    N, Hs, Ws = depth.shape
    xyz = c2m.cam2world((torch.from_numpy(depth).unsqueeze(1), torch.from_numpy(p2p).unsqueeze(1)))
    pts = xyz.permute(0, 1, 3, 4, 2).reshape(N, Hs * Ws, 3)
    xy = c2m._points_to_voxels(pts)
    mask = torch.all((pts < c2m.max_bound) & (pts >= c2m.min_bound), dim=2)
    fm = torch.from_numpy(feats) * mask.unsqueeze(1)
    vol, dens = c2m.splat_soft((xy, fm, c2m.grid_size[:2]))
    XY = xy.floor().long()
    nz = dens[:, :, 0] > 0
    np.savez_compressed(os.path.join(OUT, "splat.npz"), xyz=xyz.numpy().reshape(N, 3, -1),
                        xy=xy.numpy(), mask=mask.numpy(), XY=XY.numpy(),
                        dens=dens[:, :, 0].numpy(), vol_nz=vol.permute(0, 2, 1)[nz].numpy(),
                        nz=nz.numpy())

    # ---- splat backward (stage-2 groundwork): the reference's own autograd through splat_soft on the same case
    gq = np.random.default_rng(77)
    G = gq.standard_normal(tuple(vol.shape)).astype(np.float32)
    Gd = gq.standard_normal(tuple(dens.shape)).astype(np.float32)
    xy_g = xy.detach().clone().requires_grad_(True)
    fm_g = fm.detach().clone().requires_grad_(True)
    torch.manual_seed(0)             # splat_soft draws random indices for its weight-0 out-of-bounds votes
    vol_g, dens_g = c2m.splat_soft((xy_g, fm_g, c2m.grid_size[:2]))
    ((vol_g * torch.from_numpy(G)).sum() + (dens_g * torch.from_numpy(Gd)).sum()).backward()
    # G / Gd are regenerated from the seed by the test (default_rng(77): G [N,C,HW] then Gd [N,HW,1], float32)
    np.savez_compressed(os.path.join(OUT, "splat_bwd.npz"), xy=xy.numpy(), feats=fm.numpy(), g_seed=np.array(77),
                        grid=np.array([int(c2m.grid_size[0]), int(c2m.grid_size[1])]),
                        dfeats=fm_g.grad.numpy(), dxy=xy_g.grad.numpy())

    # ---- LiDAR raster (projection.py:64-134, build_dense_depth.py:461-463)
    H, W = 128, 240
    pc = synth.os1_scan(seed=3)[::8]
    P = synth.lidar2camrect(H, W)
    pts2, dep2 = mods["projection"].pixels_to_depth(pc, {"lidar2camrect": P}, H, W)
    img = np.zeros((H, W), np.float32)
    img[pts2[:, 1], pts2[:, 0]] = dep2
    mm = np.clip(img * 1000, 0, 65535).astype(np.uint16)
    np.savez_compressed(os.path.join(OUT, "lidar.npz"), depth_m=img, depth_mm=mm)

    # ---- depth expectation (depth_utils.py:300-313)
    import creste.models.depth as rdepth
    logits = synth.depth_logits_inputs()
    m, b = rdepth.DepthCompletion._convert_to_metric_depth(
        torch.from_numpy(logits),
        OmegaConf.create(dict(mode="UD", depth_min=300, depth_max=25600, num_bins=128)))
    np.savez_compressed(os.path.join(OUT, "depth.npz"), metric=m.numpy(), bins=b.numpy())

    # ---- expert visitation + loss value (loss_utils.py:1055-1259)
    B, Hm, Wm, T = 4, 64, 128, 50
    expert, cfs, exp_svf, reward = synth.loss_inputs(B, Hm, Wm, T)
    L = mods["loss_utils"].MaxEntIRLLoss(OmegaConf.create(cfgs["irl"]["loss"][0]))
    fov = mods["train_utils"].create_trapezoidal_fov_mask(256, 256, 70, 70, 7, 200)
    fov = fov.unsqueeze(0).repeat(B, 1, 1)
    td = {"outputs/exp_svf": torch.from_numpy(exp_svf), "inputs/traversability_label":
          torch.from_numpy(expert), "inputs/fov_mask": fov, "inputs/counterfactuals_label": cfs,
          "outputs/traversability_preds": torch.from_numpy(reward),
          "outputs/input_view": torch.zeros(B, 40, Hm, Wm)}
    ld, md = L.loss(td)
    _, cnt = L.compute_expert_visitation(torch.from_numpy(expert), 2, [Hm, Wm])
    np.savez_compressed(os.path.join(OUT, "loss.npz"), counts=cnt.numpy(), fov=fov[0].numpy(),
                        loss=np.float32(ld["maxentirl_loss"].item()),
                        mean_exp=np.float32(md["mean_expected_svf_rewards"].item()),
                        mean_svf=np.float32(md["mean_svf_rewards"].item()))

    # ---- stage-3 training step: reward FCN (train-mode BN) + MaxEntIRLLoss incl. the double
    # backward of the gradient penalty + Adam (train_traversability.py:62-103)
    irl_step_golden()
    stage1_loss_golden()
    distill_step_golden()
    bev_step_golden()
    ssc_step_golden()

    # ---- full forward, tiny image, both depth profiles (lfd.py:314-330)
    for prof in ("peaky", "soft"):
        H, W = 64, 96
        model, _ = rh.build_ref_maxentirl(image_size=(H, W))
        model.eval()
        sd = synth.seeded_state_dict(model.state_dict(), seed=0, depth_profile=prof)
        model.load_state_dict(sd)
        rgbd, p2p_t = synth.net_inputs(H, W, B=1)
        with torch.no_grad():
            out = model((rgbd, p2p_t))
        dens = out["bev_densities"][0, 0]
        nz = dens > 0
        np.savez_compressed(
            os.path.join(OUT, f"forward_{prof}_{H}x{W}.npz"),
            costmap=out["traversability_preds"].numpy(),
            depth_metric=out["depth_preds_metric"].numpy(),
            depth_bins=out["depth_preds_bins"].numpy().astype(np.int16),
            feats_sample=out["depth_preds_feats"][0, ::16].numpy(),
            dino_sample=out["dino_pe_feats"][0, 0, ::16].numpy(),
            logits_sample=out["depth_preds_logits"][0, ::16].numpy(),
            bev_nz=nz.numpy(), bev_dens_nz=dens[nz].numpy(),
            bev_feat_nz=out["bev_features"][0][:, nz].numpy(),
            input_view_sample=out["input_view"].detach()[0, ::8, ::2, ::2].numpy(),
            elevation_sample=out["elevation_preds"][0, :, ::4, ::4].numpy(),
            n_keys=np.array(len(sd)))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f"  {f}: {os.path.getsize(os.path.join(OUT, f)) / 1024:.1f} KB")


def distill_step_golden():
    """Stage-1 training step of the UNMODIFIED reference (DistillationBackbone in train mode + the three
    losses of effnet_ds2_dinov2_128.yaml + Adam, train_pefree.py:76-106, 176-181) on the seeded 64x96
    case of oracle/distill_oracle.make_case: loss values, the L2 norm of every parameter gradient, three
    full gradient tensors (stem, depth-head BN, last dino conv), samples of the outputs."""
    from . import distill_oracle as do
    ref = do.reference_step(do.make_case())
    names = sorted(ref["grads"])
    full = ["dino_head.model.6.weight", "depthcomp.depth_head.model.1.weight",
            "depthcomp.vision_backbone.model.trunk._conv_stem.weight"]
    np.savez_compressed(
        os.path.join(OUT, "distill_step.npz"),
        loss=ref["loss"], ce=ref["CrossEntropyDepth/depth/cls_loss"], sl1=ref["SmoothL1Depth/depth/reg_loss"],
        mse=ref["MSELoss/loss"], acc=ref["CrossEntropyDepth/depth/acc"],
        grad_names=np.array(names),
        grad_l2=np.array([np.sqrt((ref["grads"][n].astype(np.float64) ** 2).sum()) for n in names]),
        logits_sample=ref["logits"][:, ::16], dino_sample=ref["dino"][:, :, ::16],
        bn_running_mean=ref["params"]["depthcomp.vision_backbone.model.trunk._bn1.running_mean"],
        **{"grad::" + n: ref["grads"][n] for n in full})


def bev_step_golden():
    """Stage-2 groundwork: train-mode forward + backward of the UNMODIFIED reference BEV decoder
    (inpainting.py:70-109) on the seeded case of oracle/bev_oracle.make_case: loss, the L2 norm of every
    gradient, two full gradient tensors (the strided layer2 conv and its 1x1 downsample) and an output sample."""
    from . import bev_oracle as bo
    ref = bo.reference_step(bo.make_case())
    names = sorted(ref["grads"])
    full = ["layer2.0.conv1.weight", "layer2.0.downsample.0.weight"]
    np.savez_compressed(
        os.path.join(OUT, "bev_step.npz"), loss=ref["loss"], grad_names=np.array(names),
        grad_l2=np.array([np.sqrt((ref["grads"][n].astype(np.float64) ** 2).sum()) for n in names]),
        preds0_sample=ref["preds0"][:, ::4, ::4, ::4],
        bn1_running_mean=ref["buffers"]["bn1.running_mean"],
        **{"grad::" + n: ref["grads"][n] for n in full})


def ssc_step_golden():
    """Stage-2 graph: train-mode forward + backward of the UNMODIFIED reference TerrainNet (terrainnet.py:272-350)
    on the seeded case of oracle/ssc_oracle.make_case: the scalar, the L2 norm of every gradient, three full
    gradient tensors (z-MLP, fusion conv, the strided layer2 conv), the soft-argmax depth and a BEV sample."""
    from . import ref_harness as rh
    from . import ssc_oracle as so
    model, _ = rh.build_ref_maxentirl(image_size=(64, 96))
    template = {k[len("backbone."):]: v for k, v in model.state_dict().items() if k.startswith("backbone.")}
    ref = so.reference_step(so.make_case(template))
    names = sorted(ref["grads"])
    full = ["cam2map.z_proj.2.weight", "cam2map.vision_fusion.convs.0.weight", "bevclassifier.layer2.0.conv1.weight"]
    np.savez_compressed(
        os.path.join(OUT, "ssc_step.npz"), loss=ref["loss"], grad_names=np.array(names),
        grad_l2=np.array([np.sqrt((ref["grads"][n].astype(np.float64) ** 2).sum()) for n in names]),
        depth_metric=ref["outputs"]["depth_preds_metric"], bev_sample=ref["samples"]["bev_features"][:, ::8],
        bn_running_mean=ref["buffers"]["cam2map.vision_fusion.convs.1.running_mean"],
        **{"grad::" + n: ref["grads"][n] for n in full})


if __name__ == "__main__":
    main()
