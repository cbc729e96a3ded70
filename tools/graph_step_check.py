"""GraphedTrainStep (engine.py) against the eager stage-1 training step: same initial replica, same batches, drop-connect
off (its RNG stream is consumed differently under graph replay) -> identical losses / parameters step for step; then the
timing of both at B = 16.   python tools/graph_step_check.py [B_time]"""
import copy
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import creste_public_b200 as cb  # noqa: E402
from creste_public_b200 import configs, engine  # noqa: E402
from creste_public_b200.creste.models.blocks import effnet  # noqa: E402
from creste_public_b200.creste.train_pefree import DistillationModel  # noqa: E402
import synth_data as synth  # noqa: E402

dev = torch.device("cuda:0")
cb.set_precision("3xfp16")
H, W = 256, 480
effnet.EfficientNetB0.DROP_CONNECT = 0.0
torch.manual_seed(7)
a = DistillationModel(configs.distill_cfg((H, W))).to(dev).train()
b = copy.deepcopy(a)
batches = [{k: v.to(dev) for k, v in synth.distill_batch(2, H, W, seed=s).items()} for s in range(3)]
step = engine.GraphedTrainStep(b, batches[0])
for i, bt in enumerate(batches):
    la = float(a.training_step(bt)["loss"])
    lb = float(step(bt)["loss"])
    pa = torch.cat([p.detach().reshape(-1) for p in a.model.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in b.model.parameters()])
    ba = torch.cat([q.detach().float().reshape(-1) for q in a.model.buffers()])
    bb = torch.cat([q.detach().float().reshape(-1) for q in b.model.buffers()])
    print(f"step {i}: loss eager {la:.7f} graphed {lb:.7f}  max|dparam| {float((pa - pb).abs().max()):.3e}  "
          f"max|dbuffer| {float((ba - bb).abs().max()):.3e}", flush=True)
del a, b, step
torch.cuda.empty_cache()

Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 16
effnet.EfficientNetB0.DROP_CONNECT = 0.2
H, W = 512, 960
m = DistillationModel(configs.distill_cfg((H, W))).to(dev).train()
bt = {k: v.to(dev) for k, v in synth.distill_batch(Bt, H, W, seed=0).items()}
for _ in range(2):
    m.training_step(bt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    m.training_step(bt)
torch.cuda.synchronize()
te = (time.perf_counter() - t0) / 3
step = engine.GraphedTrainStep(m, bt)
for _ in range(2):
    step(bt)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    out = step(bt)
torch.cuda.synchronize()
tg = (time.perf_counter() - t0) / 3
print(f"B={Bt} 512x960: eager {te * 1e3:.1f} ms/step ({Bt / te:.1f} frames/s)   graphed {tg * 1e3:.1f} ms/step "
      f"({Bt / tg:.1f} frames/s)  loss {float(out['loss']):.4f}  peak mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB")
