"""CPU tests of the stage-1 (distillation) training graph (-m "not gpu").

The sm_100a kernels only run on a GPU; here their torch stand-ins (tests/torch_backend.py) are
patched in and the *graph* -- DistillationBackbone in train mode (BatchNorm batch statistics,
squeeze-excite, drop-connect, the U-Net decoder, both heads), the three stage-1 losses, the hand-written
backward of every fused node, FlatAdam -- is checked against the oracle port, which is itself pinned
bit-for-bit to the unmodified reference (build container) and to tests/golden/distill_step.npz."""
import numpy as np
import pytest
import torch

from oracle import distill_oracle as do
from oracle import ref_shims
import torch_backend as tb

HAVE_REF = ref_shims.reference_available()


def ours_step(case, device=None):
    """One DistillationModel.training_step on the mirror; same result dict as distill_oracle.port_step."""
    from creste_public_b200 import configs, engine
    from creste_public_b200.creste.train_pefree import DistillationModel
    m = DistillationModel(configs.distill_cfg(case["image_size"]))
    m.model.load_state_dict(case["state_dict"])
    m.train()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    m = m.to(dev)
    torch.manual_seed(case["seed"])
    # the reference's drop-connect uniforms: torch.rand([B,1,1,1]) from the CPU generator, block by block
    engine.drop_connect_rand = lambda B, d: torch.rand([B, 1, 1, 1]).reshape(B).to(d)
    try:
        inputs = {k: case[k].clone().to(dev) for k in ("image", "depth_label", "fimg_label")}
        outputs, loss_dict, meta, loss = None, None, None, None
        opt = m.optimizers()
        opt.zero_grad()
        outputs, loss_dict, meta, loss = m._losses(inputs)
        loss.backward()
        out = {"loss": np.float32(loss.detach().cpu()),
               "logits": outputs["depth_preds_logits"].detach().cpu().numpy().copy(),
               "dino": outputs["dino_pe_feats"].detach().cpu().numpy().copy()}
        out.update({k: np.float32(v.detach().cpu()) for k, (w, v) in loss_dict.items()})
        out.update({k: np.float32(v.detach().cpu()) for k, v in meta.items()})
        opt._gather_grads()
        out["grads"] = {k: p.grad.detach().cpu().numpy().copy() for k, p in m.model.named_parameters()}
        opt.step()
        out["params"] = {k: v.detach().cpu().numpy().copy() for k, v in m.model.state_dict().items()}
    finally:
        engine.drop_connect_rand = None
    return out


def compare(ours, ref, truth, tol_out=2e-4, tol_grad=5e-4, strict=True):
    """Parity bars of the stage-1 step.  `ref` is the fp32 oracle port (== the reference bit for bit),
    `truth` the same step in float64.  Losses <= 2e-4 relative, outputs <= tol_out of their max.

    Gradients.  The reference's own fp32 rounding noise is large on some tensors (measured against
    float64: up to 2.4e-2 of the tensor's max on the dino-head BatchNorm vectors), because the noise is
    DISCRETE: a pre-activation within ~1e-6 of zero lands on the other side of a ReLU, and in this
    768-pixel case one flipped element moves a row of the downstream weight gradient by up to 3 % and
    everything upstream by ~1e-3 (profiles/r1c_distill_parity.md).  Two criteria:
      strict   |ours - truth|_max <= 3 * |ref - truth|_max + tol_grad * |truth|_max + 1e-6  per tensor:
               as close to the exact gradient as the reference is, up to a factor 3.  Asserted for every
               tensor when `strict` (an implementation that flips the same elements as the reference
               does); otherwise the median relative L2 error over the tensors must stay <= 5e-3.
      robust   per-tensor relative L2 error against float64 <= max(3 x the reference's, 3e-2): no single
               ReLU flip reaches it, any wrong backward formula exceeds it by an order of magnitude."""
    for k in ("loss", "CrossEntropyDepth/depth/cls_loss", "SmoothL1Depth/depth/reg_loss", "MSELoss/loss"):
        np.testing.assert_allclose(ours[k], ref[k], rtol=2e-4, err_msg=k)
    np.testing.assert_allclose(ours["CrossEntropyDepth/depth/acc"], ref["CrossEntropyDepth/depth/acc"], atol=2e-3)
    for k in ("logits", "dino"):
        assert np.abs(ours[k] - ref[k]).max() <= tol_out * np.abs(ref[k]).max(), k
    assert len(ref["grads"]) >= 240
    bad, bad_l2, rels = [], [], []
    for k, g0 in ref["grads"].items():
        t = truth[k]
        d = ours["grads"][k] - t
        err, yard, tmax = np.abs(d).max(), np.abs(g0 - t).max(), np.abs(t).max()
        lim = 3 * yard + tol_grad * tmax + 1e-6
        if not err <= lim:
            bad.append((float(err / lim), k, float(err), float(yard), float(tmax)))
        tn = np.sqrt((t ** 2).sum())
        if tn > 1e-5:                                # exact-zero gradients (biases in front of a BatchNorm) aside
            rel, rel_ref = np.sqrt((d ** 2).sum()) / tn, np.sqrt(((g0 - t) ** 2).sum()) / tn
            rels.append(rel)
            if not rel <= max(3 * rel_ref, 3e-2):
                bad_l2.append((float(rel), k, float(rel_ref)))
    assert not bad_l2, sorted(bad_l2, reverse=True)[:10]
    if strict:
        assert not bad, sorted(bad, reverse=True)[:10]
    else:
        # one flipped element near the loss shifts EVERY upstream gradient by ~1e-3 of its size (117 of 244 tensors
        # were over the strict limit in the measured fp32-mode run), so the aggregate bar is on the typical tensor
        assert float(np.median(rels)) <= 5e-3, (float(np.median(rels)), len(bad), sorted(bad, reverse=True)[:5])
    for k, g1 in ours["grads"].items():          # parameters the reference leaves without a gradient
        if k not in ref["grads"]:
            assert np.abs(g1).max() == 0.0, k
    for k, p0 in ref["params"].items():
        p1 = ours["params"][k]
        if k.endswith("num_batches_tracked"):
            assert int(p0) == int(p1), k
            continue
        # Adam normalises the step to +-lr whatever the gradient's size: a parameter whose gradient is
        # rounding noise (exact zero in exact arithmetic) may move by up to lr either way
        np.testing.assert_allclose(p1, p0, rtol=2e-3, atol=1.1e-3, err_msg=k)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_port_matches_reference():
    case = do.make_case()
    ref, port = do.reference_step(case), do.port_step(case)
    for k in ("loss", "CrossEntropyDepth/depth/cls_loss", "SmoothL1Depth/depth/reg_loss", "MSELoss/loss"):
        assert ref[k] == port[k], k
    assert set(ref["grads"]) == set(port["grads"])
    for k in ref["grads"]:
        assert np.array_equal(ref["grads"][k], port["grads"][k]), k
    for k in ref["params"]:
        assert np.array_equal(ref["params"][k], port["params"][k]), k


def test_port_matches_golden(golden):
    g = golden("distill_step.npz")
    port = do.port_step(do.make_case())
    check_golden(port, g, rtol=1e-5)


def check_golden(res, g, rtol):
    for k in ("loss", "ce", "sl1", "mse"):
        name = {"loss": "loss", "ce": "CrossEntropyDepth/depth/cls_loss", "sl1": "SmoothL1Depth/depth/reg_loss",
                "mse": "MSELoss/loss"}[k]
        np.testing.assert_allclose(res[name], g[k], rtol=max(rtol, 1e-6), err_msg=k)
    names = [str(n) for n in g["grad_names"]]
    l2 = np.array([np.sqrt((res["grads"][n].astype(np.float64) ** 2).sum()) for n in names])
    big = g["grad_l2"] > 1e-4 * g["grad_l2"].max()
    np.testing.assert_allclose(l2[big], g["grad_l2"][big], rtol=max(rtol, 1e-6) * 50)
    for n in ("dino_head.model.6.weight", "depthcomp.depth_head.model.1.weight",
              "depthcomp.vision_backbone.model.trunk._conv_stem.weight"):
        ref = g["grad::" + n]
        assert np.abs(res["grads"][n] - ref).max() <= max(rtol, 1e-6) * 50 * np.abs(ref).max(), n


def test_graph_matches_port():
    case = do.make_case()
    port = do.port_step(case)
    with tb.patched():
        ours = ours_step(case)
    compare(ours, port, do.port_grads_fp64(case)[0])


def test_graph_matches_golden(golden):
    with tb.patched():
        ours = ours_step(do.make_case())
    check_golden(ours, golden("distill_step.npz"), rtol=2e-4)


def test_train_mode_no_longer_refused_but_eval_swish_frozen_is():
    from creste_public_b200 import autograd as ag
    bn = torch.nn.BatchNorm2d(8).eval()
    with pytest.raises(NotImplementedError):
        ag.bn_act(torch.zeros(1, 2, 2, 8), bn, "swish")


DDP_STEP = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import torch_backend as tb
from oracle import distill_oracle as do
from test_distill_cpu import ours_step
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# global batch of 4 frames, rank r owns frames r::world (DistributedSampler semantics); BatchNorm statistics
# and drop-connect masks stay per-rank, the product's FlatAdam all-reduces ONE flat gradient buffer.
full = do.make_case(seed=7, B=4)
mine = dict(full)
for k in ("image", "depth_label", "fimg_label"):
    mine[k] = full[k][rank::world].contiguous()
mine["seed"] = full["seed"] + rank
with tb.patched():
    out = ours_step(mine)
gs, ps = [None] * world, [None] * world
dist.all_gather_object(gs, {k: out["grads"][k] for k in ("dino_head.model.6.weight", "depthcomp.depth_head.model.0.weight")})
dist.all_gather_object(ps, out["params"])
if rank == 0:
    for k in ps[0]:
        if "running" in k or "num_batches" in k:
            continue            # per-rank BatchNorm statistics (no SyncBN in the reference)
        assert np.array_equal(ps[0][k], ps[1][k]), k          # replicas stay in lock step
    # each rank's local gradient equals the port's on its shard; the update used their mean
    ports = []
    for r in range(world):
        c = dict(full)
        for k in ("image", "depth_label", "fimg_label"):
            c[k] = full[k][r::world].contiguous()
        c["seed"] = full["seed"] + r
        ports.append(do.port_step(c))
    for k in gs[0]:
        for r in range(world):
            ref = ports[r]["grads"][k]       # robust (relative L2) criterion: see compare() on ReLU flips
            assert np.sqrt(((gs[r][k] - ref) ** 2).sum()) <= 3e-2 * np.sqrt((ref ** 2).sum()), (k, r)
        gmean = sum(p["grads"][k] for p in ports) / world
        p0 = full["state_dict"][k].clone().requires_grad_(True)
        opt = torch.optim.Adam([p0], lr=5e-4)
        p0.grad = torch.from_numpy(gmean)
        opt.step()
        d_ref = (p0.detach() - full["state_dict"][k]).numpy()
        d_ours = ps[0][k] - full["state_dict"][k].numpy()
        agree = np.mean(np.sign(d_ref) == np.sign(d_ours))
        assert agree > 0.97, (k, agree)
    print("OK")
dist.destroy_process_group()
'''


def test_stage1_data_parallel_gloo_world2(tmp_path):
    """World-size-2 stage-1 step over gloo: per-rank shards / BatchNorm statistics / drop-connect, ONE flat
    gradient all-reduce, replicas bit-identical after Adam and moving along the mean gradient."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "ddp_distill.py"
    script.write_text(DDP_STEP)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), root]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "OK" in res.stdout


def test_drop_connect_scale_matches_efficientnet():
    """engine.drop_connect_scale == efficientnet_pytorch.utils.drop_connect's per-sample multiplier."""
    from creste_public_b200 import engine
    from oracle.ref_shims import efficientnet_shim as effs
    x = torch.ones(6, 3, 2, 2)
    torch.manual_seed(11)
    ref = effs.drop_connect(x, 0.2, True)[:, 0, 0, 0]
    torch.manual_seed(11)
    engine.drop_connect_rand = lambda B, d: torch.rand([B, 1, 1, 1]).reshape(B)
    try:
        ours = engine.drop_connect_scale(6, 0.2, torch.device("cpu"))
    finally:
        engine.drop_connect_rand = None
    torch.testing.assert_close(ours, ref)
    assert set(ours.tolist()) <= {0.0, 1.25}


def test_flat_adam_views_are_128_byte_aligned():
    """Kernels read parameters (biases, BatchNorm vectors) with 16-byte vector loads straight from the flat
    buffer views: every view must start on a 128-byte boundary whatever the sizes before it."""
    from creste_public_b200.creste.train_traversability import FlatAdam
    ps = [torch.nn.Parameter(torch.randn(n)) for n in (6, 10, 1152, 3, 48)]
    vals = [p.detach().clone() for p in ps]
    with tb.patched():
        opt = FlatAdam(ps)
    base = opt.flat_p.data_ptr()
    for p, v, g in zip(ps, vals, opt.views):
        assert (p.data_ptr() - base) % 128 == 0 and (g.data_ptr() - opt.flat_g.data_ptr()) % 128 == 0
        assert torch.equal(p.detach(), v)


@pytest.mark.parametrize("C,K,R", [(112, 40, 1), (144, 200, 3), (64, 70, 1), (30, 130, 3), (6, 96, 1)])
def test_wgrad_channel_tiling(C, K, R):
    """autograd._wgrad_raw cuts wide layers into 64-channel slices for the CUDA-core kernel: the assembled
    gradient equals the one-shot torch weight gradient (kernel stand-ins; the tiling logic is what runs)."""
    from creste_public_b200 import autograd as ag
    from creste_public_b200 import ops
    saved = (ops.conv2d_wgrad, ops.chan_slice, ag.WGRAD_TC)
    ops.conv2d_wgrad = lambda x, g, R_, S_, pad: tb.wgrad_raw(x, g, R_, S_, pad[0], pad[2])
    ops.chan_slice = tb.chan_slice
    ag.WGRAD_TC = False
    try:
        g = torch.Generator().manual_seed(C + K)
        x, gy = torch.randn(2, 9, 8, C, generator=g), torch.randn(2, 9, 8, K, generator=g)   # > 64 rows: not the SE path
        got = ag._wgrad_raw(x, gy, R, R, R // 2, R // 2)
        torch.testing.assert_close(got, tb.wgrad_raw(x, gy, R, R, R // 2, R // 2), rtol=1e-5, atol=1e-5)
    finally:
        ops.conv2d_wgrad, ops.chan_slice, ag.WGRAD_TC = saved
