"""GraphedHeadStep against the eager HeadStep: same replica, same batches -> same losses / parameters step for step;
then the timing of both at B = 8, 256 x 256.   python tools/irl_graph_check.py"""
import copy
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import creste_public_b200 as cb  # noqa: E402
from creste_public_b200 import configs  # noqa: E402
from creste_public_b200.config import as_cfg  # noqa: E402
from creste_public_b200.creste.train_traversability import GraphedHeadStep, HeadStep  # noqa: E402
from creste_public_b200.creste.utils.loss_utils import LossManager  # noqa: E402
import synth_data as synth  # noqa: E402

dev = torch.device("cuda:0")
cb.set_precision("3xfp16")
Bi, Hm, Wm = 8, 256, 256
cfg = configs.irl_cfg(image_size=(64, 96), map_size=(Hm, Wm), solve_mdp=True, action_horizon=50)
keys = ("inpainting_sam_preds", "inpainting_sam_dynamic_preds", "elevation_preds")


def inputs(seed):
    feat, expert, fov, cfs = synth.head_inputs(Bi, Hm, Wm, seed=seed)
    return {k: t.to(dev) for k, t in zip(keys, feat)}, expert.to(dev), fov.to(dev), cfs


torch.manual_seed(3)
ma = cb.build_maxentirl(cfg).to(dev)
ma.backbone.eval(); ma.traversability_head.train()
mb = copy.deepcopy(ma)
ea = HeadStep(ma, LossManager(as_cfg(cfg)))
data = [inputs(s) for s in range(3)]
eb = GraphedHeadStep(mb, LossManager(as_cfg(cfg)), data[0])
for i, d in enumerate(data):
    la, oa, _ = ea(*d)
    lb, ob, _ = eb(*d)
    pa = torch.cat([p.detach().reshape(-1) for p in ma.traversability_head.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in mb.traversability_head.parameters()])
    print(f"step {i}: loss eager {float(la):.8f} graphed {float(lb):.8f}  max|dparam| {float((pa - pb).abs().max()):.3e}  "
          f"K eager {int(ma.traversability_head.last_vi_info[0])} graphed {int(mb.traversability_head.last_vi_info[0])}  "
          f"reward equal {torch.equal(oa['traversability_preds'], ob['traversability_preds'])}", flush=True)
for name, st in (("eager", ea), ("graphed", eb)):
    for _ in range(3):
        st(*data[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        st(*data[0])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {dt * 1e3:.2f} ms/step = {1 / dt:.1f} steps/s")
# where the graphed step's time goes
d = data[0]
for _ in range(3):
    eb._labels(d[1], d[2], d[3])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    eb._labels(d[1], d[2], d[3])
torch.cuda.synchronize()
print(f"label prepass (eager, host lists + syncs): {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eb.graph.replay()
e1.record()
torch.cuda.synchronize()
print(f"graph replay (GPU time): {e0.elapsed_time(e1) / 10:.2f} ms")
t0 = time.perf_counter()
for _ in range(10):
    eb.opt.step()
torch.cuda.synchronize()
print(f"optimizer step (eager): {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")
