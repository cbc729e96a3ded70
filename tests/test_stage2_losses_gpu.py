"""GPU tests of the stage-2 loss kernels (csrc/losses2.cu) against their torch restatements (tests/torch_backend.py,
themselves checked against the unmodified reference LossManager in tests/test_stage2_losses_cpu.py), and of one
TerrainNetModel.training_step on the device against the same step on the CPU stand-ins."""
import numpy as np
import pytest
import torch

import synth_data
import torch_backend as tb

pytestmark = pytest.mark.gpu


def _ops():
    from creste_public_b200 import ops
    return ops


def _close(a, b, tol=2e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-12), float((a - b).abs().max())


def test_smooth_l1_kernels(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    pred, gt = torch.randn(2, 2, 40, 50, generator=g), torch.randn(2, 2, 40, 50, generator=g) * 700
    gt[torch.rand(gt.shape, generator=g) < 0.2] = float("nan")
    gt[0, 0, 0, 0] = float("inf")
    mask = torch.rand(gt.shape, generator=g) < 0.7
    sc = torch.tensor([0.37])
    for m in (None, mask):
        acc = ops.smooth_l1(pred.cuda(), gt.cuda(), None if m is None else m.cuda(), 1e-3, 0.5)
        _close(acc, tb.smooth_l1(pred, gt, m, 1e-3, 0.5), 1e-6)
        d = ops.smooth_l1_bwd(pred.cuda(), gt.cuda(), None if m is None else m.cuda(), 1e-3, 0.5, sc.cuda())
        _close(d, tb.smooth_l1_bwd(pred, gt, m, 1e-3, 0.5, sc), 1e-6)


def test_weighted_cross_entropy_kernels(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(2, 6, 30, 40, generator=g) * 3
    labels = torch.randint(0, 6, (2, 30, 40), generator=g)
    mask = torch.rand(2, 30, 40, generator=g) < 0.6
    w = torch.rand(6, generator=g) + 0.2
    sc = torch.tensor([1.7])
    for weights, ign in ((w, -100), (None, 0), (w, 2)):
        acc = ops.ce_weighted(logits.cuda(), labels.cuda(), mask.cuda(), None if weights is None else weights.cuda(), ign)
        _close(acc, tb.ce_weighted(logits, labels, mask, weights, ign), 1e-5)
        d = ops.ce_weighted_bwd(logits.cuda(), labels.cuda(), mask.cuda(), None if weights is None else weights.cuda(),
                                ign, sc.cuda())
        _close(d, tb.ce_weighted_bwd(logits, labels, mask, weights, ign, sc), 1e-5)


@pytest.mark.parametrize("N,Na,off,D", [(300, 300, 0, 32), (130, 417, 200, 32), (64, 64, 0, 8), (1000, 2500, 1500, 32)])
def test_supcon_kernels(cuda, N, Na, off, D):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(Na, D, generator=g)
    a, nrm = tb.l2norm_rows(x)
    ya, na = ops.l2norm_rows(x.cuda())
    _close(ya, a, 1e-6)
    _close(na, nrm, 1e-6)
    dy = torch.randn(Na, D, generator=g)
    _close(ops.l2norm_rows_bwd(ya, dy.cuda(), na), tb.l2norm_rows_bwd(a, dy, nrm), 1e-5)
    f = a[off:off + N].contiguous()
    la = torch.randint(0, 7, (Na,), generator=g)
    la[5] = 99                                   # a label with no positive at all
    lf = la[off:off + N].contiguous()
    cw = None
    stats, acc = ops.supcon_fwd(f.cuda(), a.cuda(), lf.cuda(), la.cuda(), off, 0.1, cw)
    st0, acc0 = tb.supcon_fwd(f, a, lf, la, off, 0.1, cw)
    _close(acc, acc0, 1e-5)
    _close(stats[:, 2], st0[:, 2], 0)            # positive counts: exact
    lse = stats[:, 0] + torch.log(stats[:, 1])
    _close(lse, st0[:, 0] + torch.log(st0[:, 1]), 1e-5)
    sc = torch.tensor([0.01])
    df, da = ops.supcon_bwd(f.cuda(), a.cuda(), lf.cuda(), la.cuda(), off, 0.1, cw, stats, sc.cuda())
    df0, da0 = tb.supcon_bwd(f, a, lf, la, off, 0.1, cw, st0, sc)
    _close(df, df0, 5e-5)
    _close(da, da0, 5e-5)


def test_supcon_class_weights(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    a, _ = tb.l2norm_rows(torch.randn(200, 32, generator=g))
    la = torch.randint(0, 5, (200,), generator=g)
    cw = torch.rand(5, generator=g) + 0.5
    stats, acc = ops.supcon_fwd(a.cuda(), a.cuda(), la.cuda(), la.cuda(), 0, 0.2, cw.cuda())
    _close(acc, tb.supcon_fwd(a, a, la, la, 0, 0.2, cw)[1], 1e-5)
    sc = torch.tensor([1.0])
    df, da = ops.supcon_bwd(a.cuda(), a.cuda(), la.cuda(), la.cuda(), 0, 0.2, cw.cuda(), stats, sc.cuda())
    df0, da0 = tb.supcon_bwd(a, a, la, la, 0, 0.2, cw, None, sc)
    _close(df, df0, 5e-5)
    _close(da, da0, 5e-5)


def _cfg(tmp_path):
    from creste_public_b200 import configs
    wpath = str(tmp_path / "class_weights.txt")
    np.savetxt(wpath, np.array([0.55, 0.2, 0.1, 0.08, 0.05, 0.02]))
    return configs.ssc_train_cfg((64, 96), class_weights=wpath)


def test_loss_manager_matches_cpu_restatement(cuda, tmp_path):
    from test_stage2_losses_cpu import _run, _tensors
    from creste_public_b200.config import as_cfg
    from creste_public_b200.creste.utils.loss_utils import LossManager
    cfg = _cfg(tmp_path)
    data, outs = _tensors(seed=4)
    with tb.patched():
        want = _run(LossManager(as_cfg(cfg)), data, outs)
    got = _run(LossManager(as_cfg(cfg)).cuda(), {k: v.cuda() for k, v in data.items()},
               {k: v.cuda() for k, v in outs.items()})
    for k, v in want[0].items():
        np.testing.assert_allclose(got[0][k], v, rtol=5e-5, atol=1e-7, err_msg=k)
    for k, g0 in want[2].items():
        if g0 is not None:
            assert float((got[2][k].cpu() - g0).abs().max()) <= 5e-5 * float(g0.abs().max()) + 1e-9, k


def test_terrainnet_model_training_step_on_device(cuda, tmp_path):
    """One stage-2 step (six losses, backward through splat + BEV decoder + backbone, Adam) on the device against
    the same step on the CPU stand-ins; the splat's conditioning (tests/test_stage2_cpu.py) bounds the agreement."""
    import creste_public_b200 as cb
    from creste_public_b200 import engine
    from creste_public_b200.creste.train_ssc import TerrainNetModel
    cfg = _cfg(tmp_path)
    batch = synth_data.ssc_batch(2, 64, 96, seed=1)
    res = {}
    for dev in ("cpu", "cuda"):
        torch.manual_seed(0)
        m = TerrainNetModel(cfg).train()
        sd0 = {k: v.clone() for k, v in m.state_dict().items()}
        if dev == "cuda":
            m = m.cuda()
        engine.drop_connect_rand = lambda Bn, d: torch.rand([Bn, 1, 1, 1]).reshape(Bn).to(d)
        try:
            torch.manual_seed(5)
            b = {"joint": {k: (v.clone().to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}}
            if dev == "cpu":
                with tb.patched():
                    out = m.training_step((b, 0, 0))
            else:
                cb.set_precision("fp32")
                out = m.training_step((b, 0, 0))
        finally:
            engine.drop_connect_rand = None
        res[dev] = (float(out["loss"]), {k: float(v) for k, v in m.logged.items()},
                    {k: (v.detach().cpu() - sd0[k]).float() for k, v in m.state_dict().items() if v.is_floating_point()})
    np.testing.assert_allclose(res["cuda"][0], res["cpu"][0], rtol=5e-3)
    for k, v in res["cpu"][1].items():
        np.testing.assert_allclose(res["cuda"][1][k], v, rtol=2e-2, atol=1e-4, err_msg=k)
    # the Adam step moved the same parameters in the same direction
    k = "model.bevclassifier.out_heads.0.proj.weight"
    a, b_ = res["cuda"][2][k].flatten(), res["cpu"][2][k].flatten()
    assert float(torch.nn.functional.cosine_similarity(a, b_, dim=0)) > 0.95
