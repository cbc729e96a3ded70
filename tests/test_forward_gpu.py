"""GPU parity of the mirrored nn.Modules (CUDA through the C ABI) against the CPU oracle and the
golden fixtures minted from the unmodified reference.

Two kinds of check (DESIGN.md "Conditioning of the end-to-end check"):

 * stage-wise, each stage fed the ORACLE's input tensor ("given identical
   inputs"): absolute tolerances -- costmap / rewards <= 1e-4 (north_star), encoder features
   <= 1e-5 relative to max|ref|, bins and voxel indices exact;
 * end to end: the net is ill-conditioned between the depth head and the splat (a 1e-4 m depth
   change moves O(1) BEV features by O(1e-2); the bounds mask is even discontinuous), so the
   reference's OWN fp32 rounding noise already moves the costmap by 1e-3..2e-2 (the same
   reference run on two different CPUs differs by 7e-3, see test_matches_reference_golden).
   The end-to-end assertion is therefore relative to a yardstick measured in the same test: the
   largest deviation over an ensemble of {encoder evaluated exactly in fp64, four runs with the
   last encoder conv's weights perturbed by 1e-6 relative} from the plain fp32 oracle:
   |cuda - ref32| <= 3 * yardstick + 1e-4.
"""
import numpy as np
import pytest
import torch

from oracle import net_oracle as no
from oracle import synth

pytestmark = pytest.mark.gpu
H, W = 64, 96


def _model(prof, hw=(H, W)):
    import creste_public_b200 as cb
    cb.set_precision("fp32")
    m = cb.build_maxentirl(image_size=hw).eval()
    sd = synth.seeded_state_dict(m.state_dict(), 0, prof)
    m.load_state_dict(sd)
    return m.cuda(), sd


@pytest.fixture(scope="module", params=["peaky", "soft"])
def run(request, cuda):
    prof = request.param
    model, sd = _model(prof)
    rgbd, p2p = synth.net_inputs(H, W, 1)
    with torch.no_grad():
        out = model((rgbd.cuda(), p2p.cuda()))
    out = {k: v.detach().cpu() for k, v in out.items()}
    ref = no.forward(sd, rgbd, p2p)
    ens = [no.forward(sd, rgbd, p2p, encoder_fp64=True)]
    g = torch.Generator().manual_seed(0)
    k = "backbone.depthcomp.depthcomp.vision_backbone.model.conv.weight"
    for _ in range(4):
        sd2 = dict(sd)
        sd2[k] = sd[k] * (1 + 1e-6 * torch.randn(sd[k].shape, generator=g))
        ens.append(no.forward(sd2, rgbd, p2p))
    yard = {kk: max(float((e[kk].float() - ref[kk].float()).abs().max()) for e in ens)
            for kk in ref if ref[kk].is_floating_point() and not kk.startswith("_")}
    return dict(prof=prof, model=model, sd=sd, rgbd=rgbd, p2p=p2p, out=out, ref=ref, yard=yard)


def test_output_dict_surface(run):
    out, ref = run["out"], run["ref"]
    keys = [k for k in ref if not k.startswith("_")]
    assert set(out.keys()) == set(keys)
    for k in keys:
        assert tuple(out[k].shape) == tuple(ref[k].shape), k
        assert out[k].dtype == ref[k].dtype, k


def test_encoder_close_to_reference(run):
    out, ref = run["out"], run["ref"]
    for k in ("depth_preds_feats", "depth_preds_logits", "dino_pe_feats"):
        r = ref[k].view_as(out[k])
        assert float((out[k] - r).abs().max()) <= 1e-5 * float(r.abs().max()), k
    assert torch.equal(out["depth_preds_bins"], ref["depth_preds_bins"])


def test_end_to_end_within_conditioning_bound(run):
    out, ref, yard = run["out"], run["ref"], run["yard"]
    for k in ("depth_preds_metric", "bev_features", "bev_densities", "inpainting_sam_preds",
              "elevation_preds", "input_view", "traversability_preds", "traversability_preds_full"):
        r = ref[k].view_as(out[k])
        err = float((out[k] - r).abs().max())
        assert err <= 3 * yard[k] + 1e-4, f"{k}: err {err:.3e} yardstick {yard[k]:.3e}"


def test_matches_reference_golden(run, golden):
    g = golden(f"forward_{run['prof']}_{H}x{W}.npz")
    out, ref, yard = run["out"], run["ref"], run["yard"]
    # well-conditioned part: encoder features / depth bins vs the reference's own outputs
    assert np.array_equal(out["depth_preds_bins"].numpy(), g["depth_bins"].astype(np.int64))
    for key, gk in (("depth_preds_feats", "feats_sample"), ("depth_preds_logits", "logits_sample")):
        f = out[key][0, ::16].numpy()
        assert np.abs(f - g[gk]).max() <= 1e-5 * np.abs(g[gk]).max(), key
    d = out["dino_pe_feats"][0, 0, ::16].numpy()
    assert np.abs(d - g["dino_sample"]).max() <= 1e-5 * np.abs(g["dino_sample"]).max()
    # ill-conditioned part: the golden costmap was produced by the reference on another CPU; the
    # oracle on THIS CPU is itself only within the yardstick of it
    # (the yardstick is the maximum over only five 1e-6 perturbations, i.e. a coarse estimate of the
    # spread: any 1e-7-level change of summation order -- e.g. the SE pooling order of the x-blocked
    # depthwise kernel -- lands somewhere inside a few multiples of it)
    tol = 3 * yard["traversability_preds"] + 1e-4
    assert np.abs(ref["traversability_preds"].numpy() - g["costmap"]).max() <= tol
    assert np.abs(out["traversability_preds"].numpy() - g["costmap"]).max() <= 2 * tol


# ------------------------------------------------------------------ stage-wise, teacher-forced
def test_stage_splat_given_oracle_depth(run):
    """bit-exact voxel indices + <=1e-5 BEV features when the splat is fed the oracle's tensors"""
    from creste_public_b200 import ops
    model, sd, ref = run["model"], run["sd"], run["ref"]
    depth = ref["depth_preds_metric"].cuda()
    feats = ops.nchw_to_nhwc(ref["depth_preds_feats"].cuda())
    c2m = model.backbone.cam2map
    xy, z, mask = c2m.frustum(depth, run["p2p"].view(1, 4, 4).cuda())
    assert np.array_equal(xy.cpu().numpy().view(np.uint32), ref["bev_coords"].numpy().view(np.uint32))
    with torch.no_grad():
        ret, _ = c2m.forward_nhwc(depth, feats, run["p2p"].view(1, 4, 4).cuda())
    # indices: recompute with want_idx through the op and compare to the oracle's
    fused = ops.nchw_to_nhwc(ref["_fused_feats"].cuda())
    o = ops.splat_soft(xy, fused.view(1, -1, fused.shape[-1]), None, 256, 256, want_idx=True)
    assert np.array_equal(o["idx"].cpu().numpy(), ref["_splat_idx"].numpy())
    r = ref["bev_features"]
    assert float((ret["bev_features"].cpu() - r).abs().max()) <= 1e-5 * max(1.0, float(r.abs().max()))
    np.testing.assert_allclose(ret["bev_densities"].cpu().numpy(), ref["bev_densities"].numpy(),
                               atol=1e-5, rtol=1e-5)


def test_stage_bev_decoder_given_oracle_bev(run):
    from creste_public_b200 import ops
    model, ref = run["model"], run["ref"]
    with torch.no_grad():
        ret, _ = model.backbone.bevclassifier.forward_nhwc(ops.nchw_to_nhwc(ref["bev_features"].cuda()))
    for k, v in ret.items():
        r = ref[k]
        assert float((v.cpu() - r).abs().max()) <= 1e-5 * float(r.abs().max()), k


def test_stage_costmap_given_oracle_heads(run):
    """north_star tolerance: fp32 costmap / rewards within 1e-4 on identical inputs"""
    model, ref = run["model"], run["ref"]
    feat_map = {k: ref[k].cuda() for k in model.traversability_head.reward_cfg.input_keys}
    with torch.no_grad():
        o = model.traversability_head(feat_map, None, False)
    assert torch.equal(o["input_view"].cpu(), ref["input_view"])          # max-pool is exact
    assert float((o["traversability_preds"].cpu() - ref["traversability_preds"]).abs().max()) <= 1e-4
    assert float((o["traversability_preds_full"].cpu() - ref["traversability_preds_full"]).abs().max()) <= 1e-4


def test_native_612_wide_input_non_integer_upsample(cuda):
    """512x612-style odd width: the last Up stage uses the exact ratio tuple (effnet.py:64-68)."""
    hw = (64, 76)   # /32 -> 2 x 2.375: widths 76 -> 38 -> 19 -> 9 -> 4 -> 2, odd at 1/4 res
    model, sd = _model("peaky", hw)
    rgbd, p2p = synth.net_inputs(hw[0], hw[1], 1)
    with torch.no_grad():
        out = model((rgbd.cuda(), p2p.cuda()))
    ref = no.forward(sd, rgbd, p2p)
    r = ref["depth_preds_feats"]
    assert tuple(out["depth_preds_feats"].shape) == tuple(r.shape)
    assert float((out["depth_preds_feats"].cpu() - r).abs().max()) <= 1e-5 * float(r.abs().max())


def test_batch_split_invariance(cuda):
    """frames are independent units (the data-parallel sharding axis): B=2 equals two B=1 runs"""
    model, sd = _model("peaky")
    rgbd, p2p = synth.net_inputs(H, W, 2)
    with torch.no_grad():
        both = model((rgbd.cuda(), p2p.cuda()))["traversability_preds"].cpu()
        one = torch.cat([model((rgbd[i:i + 1].cuda(), p2p[i:i + 1].cuda()))["traversability_preds"].cpu()
                         for i in range(2)])
    # everything up to the splat is bit-reproducible and batch-independent (asserted on the
    # encoder output below); the splat's atomic accumulation order is not, and the BEV decoder
    # amplifies that 1e-7-relative noise to <= 1e-4 on the costmap (the reference's CUDA
    # scatter_add_ has the same property)
    assert float((both - one).abs().max()) <= 1e-4
    with torch.no_grad():
        f2 = model((rgbd.cuda(), p2p.cuda()))["depth_preds_feats"].cpu()
        f1 = model((rgbd[1:2].cuda(), p2p[1:2].cuda()))["depth_preds_feats"].cpu()
    assert torch.equal(f2[1:2], f1)


@pytest.mark.parametrize("mode", ["3xtf32", "3xfp16"])
def test_forward_tensor_core_modes(cuda, mode):
    """Whole forward in the fp32-faithful tensor-core modes (tcgen05 3xTF32 / 3xFP16 splits):
    encoder features within 5e-5 * max of the fp32 oracle, depth bins identical on the soft
    profile, costmap within the conditioning yardstick used by the fp32 end-to-end test
    (|cuda - ref32| <= 3 * yardstick + 1e-4; the full-size stage-wise checks of the benchmarked mode are in
    tests/test_forward_full_gpu.py)."""
    import creste_public_b200 as cb
    model, sd = _model("soft")
    rgbd, p2p = synth.net_inputs(H, W, 1)
    ref = no.forward(sd, rgbd, p2p)
    ens = [no.forward(sd, rgbd, p2p, encoder_fp64=True)]
    g = torch.Generator().manual_seed(0)
    kw = "backbone.depthcomp.depthcomp.vision_backbone.model.conv.weight"
    for _ in range(4):
        sd2 = dict(sd)
        sd2[kw] = sd[kw] * (1 + 1e-6 * torch.randn(sd[kw].shape, generator=g))
        ens.append(no.forward(sd2, rgbd, p2p))
    yard = max(float((e["traversability_preds"] - ref["traversability_preds"]).abs().max()) for e in ens)
    cb.set_precision(mode)
    try:
        with torch.no_grad():
            out = model((rgbd.cuda(), p2p.cuda()))
    finally:
        cb.set_precision("fp32")
    r = ref["depth_preds_feats"]
    assert float((out["depth_preds_feats"].cpu() - r).abs().max()) <= 5e-5 * float(r.abs().max())
    agree = float((out["depth_preds_bins"].cpu() == ref["depth_preds_bins"]).float().mean())
    assert agree >= 0.999, agree
    # the costmap is conditioning-limited (perturbing one encoder layer by 1e-6 moves it
    # by 2e-3..2e-2): held to the yardstick measured above, exactly as the fp32 end-to-end test
    err = float((out["traversability_preds"].cpu() - ref["traversability_preds"]).abs().max())
    assert err <= 3 * yard + 1e-4, (err, yard)


def test_training_mode_follows_module_flag(cuda):
    """model.train() switches the whole forward to batch-statistics BatchNorm (+ running-stat updates) like the
    reference's modules do; model.eval() afterwards restores the inference engine and sees the updated statistics."""
    model, _ = _model("peaky")
    rgbd, p2p = synth.net_inputs(H, W, 2)
    with torch.no_grad():
        ev = model((rgbd.cuda(), p2p.cuda()))["traversability_preds"].clone()
    model.train()
    bn = model.backbone.bevclassifier.bn1
    before = bn.running_mean.clone()
    with torch.no_grad():
        tr = model((rgbd.cuda(), p2p.cuda()))["traversability_preds"].clone()
    assert not torch.equal(before, bn.running_mean) and int(bn.num_batches_tracked) == 1
    assert float((tr - ev).abs().max()) > 1e-4                      # batch statistics != running statistics
    model.eval()
    with torch.no_grad():
        ev2 = model((rgbd.cuda(), p2p.cuda()))["traversability_preds"]
    assert float((ev2 - ev).abs().max()) > 1e-6                     # the folded BatchNorm factors were rebuilt


def test_irl_forward_solve_mdp(cuda):
    """MaxEntIRL.forward with solve_mdp=True: VI + SVF on the model's own reward map vs the
    C oracle fed the same reward."""
    import creste_public_b200 as cb
    from oracle import c_oracle as co
    cb.set_precision("fp32")
    m = cb.build_maxentirl(image_size=(H, W), solve_mdp=True).eval()
    sd = synth.seeded_state_dict(m.state_dict(), 0, "peaky")
    m.load_state_dict(sd)
    m = m.cuda()
    rgbd, p2p = synth.net_inputs(H, W, 2)
    expert = torch.from_numpy(synth.expert_poses(2, 50, 256, 256, seed=5))
    with torch.no_grad():
        out = m((rgbd.cuda(), p2p.cuda(), expert.cuda()))
    r = out["traversability_preds"].cpu().numpy()
    v0, q0, pi0, K0 = co.vi_solve(r)
    assert int(m.traversability_head.last_vi_info[0]) == K0
    assert np.array_equal(out["value_estimate"].cpu().numpy()[:, 0].view(np.uint32), v0.view(np.uint32))
    np.testing.assert_allclose(out["policy"].cpu().numpy(), pi0, atol=1e-6)
    s0, st0, g0 = co.svf(out["policy"].cpu().numpy(), expert[:, :, :2, 2].numpy(),
                         m.fov_mask[0, 0].numpy(), 50, 2, True, 0.005, False)
    assert np.array_equal(out["state_preds"].cpu().numpy(), st0)
    np.testing.assert_allclose(out["exp_svf"].cpu().numpy(), s0, atol=2e-5, rtol=1e-5)


def test_cuda_graph_replay_equals_eager(cuda):
    """engine.GraphedForward: the captured forward replays bit-identically (up to the splat's
    atomic order) on new inputs of the same shape."""
    from creste_public_b200.engine import GraphedForward
    model, sd = _model("peaky")
    a, p2p = synth.net_inputs(H, W, 1, seed=0)
    b, _ = synth.net_inputs(H, W, 1, seed=1)
    g = GraphedForward(lambda x, p: model((x, p)), (a.cuda(), p2p.cuda()))
    with torch.no_grad():
        eager = {k: v.clone() for k, v in model((b.cuda(), p2p.cuda())).items() if torch.is_tensor(v)}
    out = g(b.cuda(), p2p.cuda())
    torch.cuda.synchronize()
    assert torch.equal(out["depth_preds_feats"], eager["depth_preds_feats"])
    assert torch.equal(out["depth_preds_bins"], eager["depth_preds_bins"])
    assert float((out["traversability_preds"] - eager["traversability_preds"]).abs().max()) <= 1e-4
