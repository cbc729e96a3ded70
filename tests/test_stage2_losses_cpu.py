"""CPU tests (-m "not gpu") of the stage-2 losses and the TerrainNetModel step: the mirror's LossManager (kernels
replaced by their torch stand-ins, tests/torch_backend.py) against the UNMODIFIED reference LossManager on the same
tensors -- values and gradients -- and the all-gather of SupPixelConLoss across two gloo ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shims
import synth_data
import torch_backend as tb

HAVE_REF = ref_shims.reference_available()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, G = 2, 256


def _cfg(tmp_path):
    from creste_public_b200 import configs
    wpath = str(tmp_path / "class_weights.txt")
    np.savetxt(wpath, np.array([0.55, 0.2, 0.1, 0.08, 0.05, 0.02]))
    return configs.ssc_train_cfg((64, 96), class_weights=wpath)


def _tensors(seed=0):
    g = torch.Generator().manual_seed(seed)
    data = synth_data.ssc_batch(B, 64, 96, seed=seed, G=G)
    outs = {"inpainting_sam_preds": torch.randn(B, 32, G, G, generator=g),
            "inpainting_sam_dynamic_preds": torch.randn(B, 6, G, G, generator=g) * 2,
            "elevation_preds": torch.randn(B, 2, G, G, generator=g) * 0.3,
            "dino_pe_feats": torch.randn(B, 1, 128, 16, 24, generator=g),
            "depth_preds_logits": torch.randn(B, 128, 16, 24, generator=g),
            "depth_preds_metric": torch.rand(B, 16, 24, generator=g) * 25}
    return data, outs


def _run(manager, data, outs, seed=7):
    data = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()}
    outs = {k: v.clone().requires_grad_(True) for k, v in outs.items()}
    td = {f"inputs/{k}": v for k, v in data.items()}
    td.update({f"outputs/{k}": v for k, v in outs.items()})
    td["task"] = "joint"
    torch.manual_seed(seed)                        # the per-class random subsampling of SupPixelConLoss
    loss_dict, meta = manager(td)
    total = sum(w * v for w, v in loss_dict.values())
    total.backward()
    return ({k: float(w) * float(v) for k, (w, v) in loss_dict.items()}, {k: float(v) for k, v in meta.items()},
            {k: (v.grad.clone() if v.grad is not None else None) for k, v in outs.items()}, float(total))


@pytest.fixture(scope="module")
def gloo1(tmp_path_factory):
    """The reference's MultiPosConLoss calls torch.distributed.nn.all_gather unconditionally: world size 1."""
    import torch.distributed as dist
    f = tmp_path_factory.mktemp("pg") / "store"
    dist.init_process_group("gloo", init_method=f"file://{f}", rank=0, world_size=1)
    yield
    dist.destroy_process_group()


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_stage2_losses_match_reference(tmp_path, gloo1):
    from oracle import ref_harness as rh
    from creste_public_b200.config import as_cfg
    from creste_public_b200.creste.utils.loss_utils import LossManager
    cfg = _cfg(tmp_path)
    mods = rh.ref_modules()
    from omegaconf import OmegaConf
    ref_mgr = mods["loss_utils"].LossManager(OmegaConf.create(cfg))
    data, outs = _tensors()
    want = _run(ref_mgr, data, outs)
    with tb.patched():
        got = _run(LossManager(as_cfg(cfg)), data, outs)
    assert set(got[0]) == set(want[0]) and len(want[0]) == 7        # SupPixelConLoss contributes two entries
    for k, v in want[0].items():
        np.testing.assert_allclose(got[0][k], v, rtol=2e-5, atol=1e-7, err_msg=k)
    for k, v in want[1].items():
        np.testing.assert_allclose(got[1][k], v, rtol=1e-5, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(got[3], want[3], rtol=2e-5)
    for k, g0 in want[2].items():
        assert (g0 is None) == (got[2][k] is None), k
        if g0 is not None:
            assert float((got[2][k] - g0).abs().max()) <= 2e-5 * float(g0.abs().max()) + 1e-9, k


def test_supcon_mask_cache_quirk():
    """The reference rebuilds its positives mask only when the local row count changes (supcon_loss.py:87-99)."""
    from creste_public_b200.creste.models.losses.supcon_loss import MultiPosConLoss
    torch.manual_seed(0)
    f = torch.randn(12, 8)
    l1, l2 = torch.arange(12) % 3, torch.arange(12) % 4
    with tb.patched():
        m = MultiPosConLoss(0.1)
        a = float(m({"feats": f, "labels": l1})["loss"])
        b = float(m({"feats": f, "labels": l2})["loss"])          # same N: the first call's labels are reused
        c = float(MultiPosConLoss(0.1)({"feats": f, "labels": l2})["loss"])
    assert a == b and abs(b - c) > 1e-4


GATHER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch_backend as tb
from creste_public_b200.creste.models.losses.supcon_loss import MultiPosConLoss
dist.init_process_group("gloo")
rank = dist.get_rank()
g = torch.Generator().manual_seed(5)
n = [10, 7]                                           # ragged: ranks hold different row counts
feats = [torch.randn(k, 8, generator=g) for k in n]
labels = [torch.randint(0, 3, (k,), generator=g) for k in n]
x = feats[rank].clone().requires_grad_(True)
with tb.patched():
    loss = MultiPosConLoss(0.1)({"feats": x, "labels": labels[rank]})["loss"]
    loss.backward()
# single-process restatement: both ranks' losses over the concatenated rows, autograd through everything
xs = [f.clone().requires_grad_(True) for f in feats]
fa = torch.cat([torch.nn.functional.normalize(t, dim=-1) for t in xs])
la = torch.cat(labels)
tot, mine = 0.0, None
for r in range(2):
    off = sum(n[:r])
    li, _, _ = tb._supcon_rows(fa[off:off + n[r]], fa, labels[r], la, off, 0.1, None)
    tot = tot + li.mean()
    if r == rank:
        mine = li.mean()
tot.backward()
assert abs(float(loss) - float(mine)) < 1e-5, (float(loss), float(mine))
assert float((x.grad - xs[rank].grad).abs().max()) < 1e-5 * float(xs[rank].grad.abs().max()) + 1e-7
dist.barrier()
if rank == 0:
    print("OK")
dist.destroy_process_group()
'''


def test_supcon_all_gather_with_gradient_gloo_world2(tmp_path):
    script = tmp_path / "gather.py"
    script.write_text(GATHER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29641", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "OK" in res.stdout


def test_terrainnet_model_training_step_runs_and_updates(tmp_path):
    """TerrainNetModel.training_step end to end on the CPU stand-ins: six losses, backward through the whole
    network, FlatAdam step; the backbone freeze schedule rebuilds the optimiser over the trainable set."""
    from creste_public_b200.creste.train_ssc import TerrainNetModel
    cfg = _cfg(tmp_path)
    cfg["freeze_backbone_epochs"] = 1
    torch.manual_seed(0)
    m = TerrainNetModel(cfg).train()
    batch = {"joint": synth_data.ssc_batch(B, 64, 96, seed=1, G=G)}
    w0 = m.model.bevclassifier.conv1.weight.detach().clone()
    e0 = m.model.depthcomp.depthcomp.vision_backbone.model.conv.weight.detach().clone()
    with tb.patched():
        m.on_train_epoch_start()                    # epoch 0 < freeze_backbone_epochs: backbone frozen
        assert m.backbone_frozen and not any(p.requires_grad for p in m.model.depthcomp.parameters())
        out = m.training_step((batch, 0, 0))
        n_frozen = len(m.optimizers().params)
        assert torch.isfinite(out["loss"]) and len(m.logged) >= 9
        assert not torch.equal(w0, m.model.bevclassifier.conv1.weight) and torch.equal(
            e0, m.model.depthcomp.depthcomp.vision_backbone.model.conv.weight)
        m.on_train_epoch_end()
        m.on_train_epoch_start()                    # epoch 1: unfrozen, optimiser rebuilt over all parameters
        assert not m.backbone_frozen
        m.training_step(({"joint": synth_data.ssc_batch(B, 64, 96, seed=2, G=G)}, 0, 0))
        assert len(m.optimizers().params) > n_frozen
        assert not torch.equal(e0, m.model.depthcomp.depthcomp.vision_backbone.model.conv.weight)
        v = m.validation_step((batch, 0, 0))
    assert torch.isfinite(v["loss"]) and "val/loss" in m.logged
