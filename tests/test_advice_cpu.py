"""CPU tests (-m "not gpu") of host-side rules that keep the mirror's behaviour equal to the reference's
around training: pack-cache invalidation after raw-pointer parameter updates, checkpoint loading rules of
TerrainNet.load_weights (reference creste/models/terrainnet.py:111-261), DDP-style initial-state and buffer
broadcast (gloo, world size 2), validation under no_grad."""
import os
import subprocess
import sys

import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pack_cache_sees_raw_pointer_writes():
    """A kernel writing through data_ptr() does not advance torch's version counter (the fused Adam on the flat
    buffer, the BatchNorm running-statistics update); engine.mark_written must invalidate the packs."""
    from creste_public_b200 import engine
    p = nn.Parameter(torch.zeros(4))
    flat = torch.zeros(8)
    p.data = flat[:4]
    cache, builds = engine.PackCache(), []

    def build():
        builds.append(1)
        return p.detach().clone()
    cache.get("w", [p], build)
    cache.get("w", [p], build)
    assert len(builds) == 1
    flat.add_(1)                                  # what creste_adam_step does: no version bump on p
    assert torch.equal(cache.get("w", [p], build), torch.zeros(4)) and len(builds) == 1      # stale!
    engine.mark_written([p, None])
    assert torch.equal(cache.get("w", [p], build), torch.ones(4)) and len(builds) == 2


def test_flat_adam_step_and_bn_update_invalidate_packs():
    import torch_backend as tb
    from creste_public_b200 import engine
    from creste_public_b200.creste.train_traversability import FlatAdam
    conv = nn.Conv2d(4, 4, 1)
    fc = engine.FusedConv(conv, None)
    with tb.patched():
        opt = FlatAdam(conv.parameters(), lr=1e-1)
        v0 = engine._ver(conv.weight, conv.bias)
        opt.zero_grad()
        (conv.weight.sum() + conv.bias.sum()).backward()
        opt.step()
    assert engine._ver(conv.weight, conv.bias) != v0
    assert fc.conv is conv


def _terrainnet():
    import creste_public_b200 as cb
    return cb.build_terrainnet(image_size=(64, 96))


def test_terrainnet_load_weights_rules(tmp_path):
    """'loss.*' keys of a Lightning stage-2 checkpoint are dropped in every strict mode; strict_unfreezesplat loads
    non-strictly and leaves exactly the cam2map parameters trainable; ft_* modes follow the name rules."""
    m = _terrainnet()
    sd = {"model." + k: v.clone() for k, v in m.state_dict().items()}
    sd["loss.losses.0.some_buffer"] = torch.zeros(3)
    k0 = "model.bevclassifier.conv1.weight"
    sd[k0] = torch.full_like(sd[k0], 0.25)
    path = str(tmp_path / "ckpt.pt")
    torch.save({"state_dict": sd}, path)

    for setting in ("strict", "strict_freeze"):
        n = _terrainnet()
        n.load_setting = setting
        n.load_weights(path)                                  # must not raise on the loss.* key
        assert float(n.bevclassifier.conv1.weight.mean()) == 0.25
        want = setting == "strict"
        assert all(p.requires_grad == want for p in n.parameters())

    # strict_unfreezesplat: non-strict (a missing key is tolerated), cam2map trainable, the rest frozen
    sd2 = dict(sd)
    del sd2["model.bevclassifier.out_heads.0.proj.bias"]
    torch.save({"state_dict": sd2}, path)
    n = _terrainnet()
    n.load_setting = "strict_unfreezesplat"
    n.load_weights(path)
    for name, p in n.named_parameters():
        assert p.requires_grad == ("cam2map." in name), name
    n = _terrainnet()
    n.load_setting = "strict"
    with pytest.raises(RuntimeError):
        n.load_weights(path)

    torch.save({"state_dict": sd}, path)
    n = _terrainnet()
    n.load_setting = "ft_decoders_all"
    n.load_weights(path)
    assert float(n.bevclassifier.conv1.weight.mean()) == 0.25
    for name, p in n.named_parameters():
        assert p.requires_grad == ("bevclassifier.out_heads" in name), name
    n = _terrainnet()
    n.load_setting = "ft_decoders_partial"
    n.load_weights(path)
    for name, p in n.named_parameters():
        last = "bevclassifier.out_heads" in name and ("up2" in name or "proj" in name)
        assert p.requires_grad == last, name
    n = _terrainnet()
    n.load_setting = "ft_semantic_head"
    n.load_weights(path)
    assert not any(p.requires_grad for p in n.parameters())   # no 1-channel head in the shipped config


def test_stage1_key_surgery_in_terrainnet_load(tmp_path):
    """Stage-1 checkpoints store depthcomp.* / dino_head.* one level up (terrainnet.py:125-140)."""
    m = _terrainnet()
    sd = {}
    for k, v in m.state_dict().items():
        if k.startswith("depthcomp.depthcomp."):
            k = k.replace("depthcomp.depthcomp.", "depthcomp.", 1)
        elif k.startswith("depthcomp.dino_head."):
            k = k.replace("depthcomp.dino_head.", "dino_head.", 1)
        sd["model." + k] = v.clone()
    path = str(tmp_path / "s1.pt")
    torch.save({"state_dict": sd}, path)
    n = _terrainnet()
    n.load_setting = "strict"
    n.load_weights(path)


def test_maxentirl_refuses_trainable_splat():
    import creste_public_b200 as cb
    from creste_public_b200 import configs
    cfg = configs.irl_cfg(image_size=(64, 96))
    cfg["vision_backbone"]["load_setting"] = "strict_unfreezesplat"
    with pytest.raises(NotImplementedError, match="strict_unfreezesplat"):
        cb.build_maxentirl(cfg)


def test_depth_loss_checks_num_bins():
    from creste_public_b200.creste.utils.loss_utils import _Stage1DepthValues
    td = {"outputs/depth_preds_logits": torch.zeros(1, 64, 4, 4), "outputs/depth_preds_bins": torch.zeros(1, 4, 4).long(),
          "inputs/depth_label": torch.zeros(1, 1, 4, 4)}
    with pytest.raises(ValueError, match="num_bins"):
        _Stage1DepthValues.get(td, {"mode": "UD", "num_bins": 128, "depth_min": 300, "depth_max": 25600}, 0.5)


BCAST = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch_backend as tb
from torch import nn
from creste_public_b200.creste.train_traversability import FlatAdam, broadcast_buffers
dist.init_process_group("gloo")
rank = dist.get_rank()
torch.manual_seed(100 + rank)                      # per-rank RNG: replicas start from DIFFERENT weights
net = nn.Sequential(nn.Conv2d(3, 5, 3), nn.BatchNorm2d(5), nn.Conv2d(5, 2, 1))
net[1].running_mean.normal_(); net[1].running_var.uniform_(0.5, 2.0); net[1].num_batches_tracked += 3 + rank
with tb.patched():
    opt = FlatAdam(net.parameters(), lr=1e-3)
broadcast_buffers(net)
state = {k: v.clone() for k, v in net.state_dict().items()}
got = [None, None]
dist.all_gather_object(got, state)
if rank == 0:
    for k in got[0]:
        assert torch.equal(got[0][k], got[1][k]), k
    assert int(got[1]["1.num_batches_tracked"]) == 3
    print("OK")
dist.destroy_process_group()
'''


def test_flat_adam_broadcasts_initial_state_gloo_world2(tmp_path):
    """DDP semantics (the reference trains under Lightning's DDPStrategy): rank 0's parameters at wrap time, rank
    0's buffers before every forward."""
    script = tmp_path / "bcast.py"
    script.write_text(BCAST)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "OK" in res.stdout
