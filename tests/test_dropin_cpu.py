"""Drop-in boundary (-m "not gpu"): the mirror is OVERLAID on the reference checkout
(creste_public_b200.install_as_creste(reference_root)) and the reference's OWN train scripts -- unmodified
creste/train_traversability.py, creste/train_pefree.py, creste/train_ssc.py -- are imported over it and their
LightningModule.training_step / validation_step / configure_optimizers are driven with configs composed from the
reference's YAML files.  The kernels only run on a GPU, so their torch stand-ins (tests/torch_backend.py) are patched
in; what is under test is the SURFACE: import paths, class lookup by config strings, constructor / forward
signatures, output-dict keys consumed by the reference's LossManager wiring, parameters the reference's Adam picks
up, and that a step changes them.  Each case runs in its own interpreter (the overlay rewires sys.modules)."""
import os
import subprocess
import sys

import pytest

from oracle import ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")

PRELUDE = r'''
import importlib.util, os, sys
import numpy as np, torch
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_shims, ref_harness as rh
ref_shims.install(); ref_shims.install_lightning()
import creste_public_b200 as cb
cb.install_as_creste(ref_shims.REFERENCE_ROOT)
import torch_backend as tb, synth_data
from omegaconf import OmegaConf

def load_script(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(ref_shims.REFERENCE_ROOT, "creste", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

def mirror_file(obj):
    f = sys.modules[type(obj).__module__].__file__
    assert f.startswith(os.path.join(ROOT, "creste_public_b200")), f
'''

IRL = PRELUDE + r'''
import creste.models.lfd, creste.utils.visualization, creste.datasets.coda_utils
assert creste.models.lfd.__file__.startswith(ROOT) and creste.utils.visualization.__file__.startswith(ref_shims.REFERENCE_ROOT)
mod = load_script("train_traversability")
cfg = rh.compose_cfgs(image_size=(64, 96))["irl"]
cfg["solve_mdp"] = True
m = mod.MaxEntIRLModel(OmegaConf.create(cfg))
mirror_file(m.model); mirror_file(m.loss); mirror_file(m.model.backbone.cam2map)
B = 2
rgbd, p2p = synth_data.distill_batch(B, 64, 96, seed=0)["image"], torch.from_numpy(synth_data.make_p2p(64, 96)).view(1, 1, 4, 4).repeat(B, 1, 1, 1)
expert = torch.from_numpy(synth_data.expert_poses(B, 50, 256, 256, seed=3))
fov = torch.from_numpy(np.asarray(synth_data.trapezoid_fov_mask(256, 256, 70, 70, 7, 200), bool)).unsqueeze(0).repeat(B, 1, 1)
data = {"image": rgbd, "p2p": p2p, "traversability_label": expert, "fov_mask": fov,
        "counterfactuals_label": synth_data.counterfactuals(expert.numpy())}
head = m.model.traversability_head
w0 = head.r.prepool[0].conv.weight.detach().clone()
b0 = m.model.backbone.bevclassifier.conv1.weight.detach().clone()
with tb.patched():
    out = m.training_step(({"joint": data}, 0, 0))
    opt = m.optimizers()
assert type(opt).__name__ == "Adam" and torch.isfinite(out["loss"])
n_opt = sum(p.numel() for g in opt.param_groups for p in g["params"])
assert n_opt == sum(p.numel() for p in m.model.parameters() if p.requires_grad)
assert not torch.equal(w0, head.r.prepool[0].conv.weight)            # the reference's Adam stepped the mirror's head
assert "train/loss" in m.logged and any(k.startswith("train/MaxEntIRLLoss") for k in m.logged)
with tb.patched(), torch.no_grad():
    m.log_img_outputs = lambda *a, **k: None                          # visualisation (mocked matplotlib): not under test
    v = m.validation_step(({"joint": data}, 0, 0))
assert torch.isfinite(v["loss"])
print("OK", float(out["loss"]))
'''

DISTILL = PRELUDE + r'''
mod = load_script("train_pefree")
cfg = rh.compose_cfgs(image_size=(64, 96))["distill"]
m = mod.DistillationModel(OmegaConf.create(cfg))
mirror_file(m.model); mirror_file(m.loss)
batch = synth_data.distill_batch(2, 64, 96, seed=0)
batch["p2p"] = torch.from_numpy(synth_data.make_p2p(64, 96)).view(1, 1, 4, 4).repeat(2, 1, 1, 1)
w0 = m.model.depthcomp.depth_head.model[0].weight.detach().clone()
with tb.patched():
    opts, scheds = m.configure_optimizers()
    opts[0].zero_grad()
    out = m.training_step(batch)                                     # Lightning's automatic optimisation: caller steps
    loss = out["loss"] if isinstance(out, dict) else out
    loss.backward()
    opts[0].step()
assert torch.isfinite(loss) and not torch.equal(w0, m.model.depthcomp.depth_head.model[0].weight)
assert any(k.startswith("train/") for k in m.logged)
print("OK", float(loss))
'''

SSC = PRELUDE + r'''
import torch.distributed as dist
dist.init_process_group("gloo", init_method="file://" + sys.argv[2], rank=0, world_size=1)   # MultiPosConLoss gathers
mod = load_script("train_ssc")
cfg = rh.compose_cfgs(image_size=(64, 96))["ssc"]
wfile = sys.argv[2] + ".weights.txt"
np.savetxt(wfile, np.array([0.55, 0.2, 0.1, 0.08, 0.05, 0.02]))
for l in cfg["loss"]:
    if "class_weights" in l:
        l["class_weights"] = wfile
m = mod.TerrainNetModel(OmegaConf.create(cfg))
mirror_file(m.model); mirror_file(m.loss); mirror_file(m.model.bevclassifier)
batch = synth_data.ssc_batch(2, 64, 96, seed=1)
w0 = m.model.bevclassifier.out_heads[0].proj.weight.detach().clone()
z0 = m.model.cam2map.z_proj[0].weight.detach().clone()
with tb.patched():
    opts, scheds = m.configure_optimizers()
    opts[0].zero_grad()
    out = m.training_step(({"joint": batch}, 0, 0))
    out["loss"].backward()
    opts[0].step()
assert torch.isfinite(out["loss"])
assert not torch.equal(w0, m.model.bevclassifier.out_heads[0].proj.weight)
assert not torch.equal(z0, m.model.cam2map.z_proj[0].weight)          # the splat's backward reached the z-MLP
assert any("SupPixelConLoss" in k for k in m.logged)
dist.destroy_process_group()
print("OK", float(out["loss"]))
'''


def _run(tmp_path, name, body):
    script = tmp_path / f"{name}.py"
    script.write_text(body)
    res = subprocess.run([sys.executable, str(script), ROOT, str(tmp_path / "pg")], capture_output=True, text=True,
                         timeout=1200)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
    assert "OK" in res.stdout


def test_reference_train_traversability_drives_the_mirror(tmp_path):
    _run(tmp_path, "irl", IRL)


def test_reference_train_pefree_drives_the_mirror(tmp_path):
    _run(tmp_path, "distill", DISTILL)


def test_reference_train_ssc_drives_the_mirror(tmp_path):
    _run(tmp_path, "ssc", SSC)


def test_install_as_creste_without_reference_tree():
    """Without a checkout the mirror alone answers `import creste...`, and a name it does not provide says so."""
    code = ("import sys; sys.path.insert(0, %r); import creste_public_b200 as cb; cb.install_as_creste();"
            "import creste.models.lfd as l, creste.utils.depth_utils as d, creste.train_ssc as t;"
            "from creste.utils.loss_utils import LossManager, SupPixelConLoss;"
            "assert l.__file__.startswith(%r)\n"
            "try:\n    from creste.utils.utils import make_labels_contiguous_vectorized\n"
            "except ImportError as e:\n    print('OK', 'not overlaid' in str(e) or 'cannot import' in str(e))\n") % (ROOT, ROOT)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK True" in res.stdout, res.stdout + res.stderr
