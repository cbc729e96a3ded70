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


def compare(ours, ref, tol_out=2e-4, tol_grad=2e-3, tol_par=2e-3):
    """Parity bars of the stage-1 step (fp32; reductions are ordered differently from oneDNN's):
    losses <= 2e-4 relative, outputs <= tol_out of their max, every gradient tensor <= tol_grad of
    its own max (+ a 1e-6 floor: bias gradients in front of a train-mode BatchNorm are exact zeros up to
    rounding), post-Adam parameters / running statistics <= tol_par absolute-relative."""
    for k in ("loss", "CrossEntropyDepth/depth/cls_loss", "SmoothL1Depth/depth/reg_loss", "MSELoss/loss"):
        np.testing.assert_allclose(ours[k], ref[k], rtol=2e-4, err_msg=k)
    np.testing.assert_allclose(ours["CrossEntropyDepth/depth/acc"], ref["CrossEntropyDepth/depth/acc"], atol=2e-3)
    for k in ("logits", "dino"):
        assert np.abs(ours[k] - ref[k]).max() <= tol_out * np.abs(ref[k]).max(), k
    live = 0
    for k, g0 in ref["grads"].items():
        g1 = ours["grads"][k]
        err = np.abs(g0 - g1).max()
        assert err <= tol_grad * np.abs(g0).max() + 1e-6, (k, err, np.abs(g0).max())
        live += 1
    assert live >= 240
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
        np.testing.assert_allclose(p1, p0, rtol=tol_par, atol=1.1e-3, err_msg=k)


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
    compare(ours, port)


def test_graph_matches_golden(golden):
    with tb.patched():
        ours = ours_step(do.make_case())
    check_golden(ours, golden("distill_step.npz"), rtol=2e-4)


def test_train_mode_no_longer_refused_but_eval_swish_frozen_is():
    from creste_public_b200 import autograd as ag
    bn = torch.nn.BatchNorm2d(8).eval()
    with pytest.raises(NotImplementedError):
        ag.bn_act(torch.zeros(1, 2, 2, 8), bn, "swish")
