"""A few stage-2 (train_ssc.py mirror) training steps at 512x960 for profiling.
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv python tools/stage2_step.py [B] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb  # noqa: E402
from creste_public_b200 import _lib, configs  # noqa: E402
from creste_public_b200.creste.train_ssc import TerrainNetModel  # noqa: E402
import synth_data as synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cb.set_precision("3xfp16")
torch.manual_seed(0)
m = TerrainNetModel(configs.ssc_train_cfg((512, 960))).cuda().train()
batch = {"joint": {k: v.cuda() for k, v in synth.ssc_batch(B, 512, 960, seed=0).items()}}
for i in range(steps):
    torch.cuda.synchronize()
    n0, t0 = _lib.lib().creste_launch_count(), time.perf_counter()
    out = m.training_step((batch, 0, 0))
    torch.cuda.synchronize()
    print(f"step {i}: {1e3 * (time.perf_counter() - t0):.1f} ms, {_lib.lib().creste_launch_count() - n0} launches, "
          f"loss {float(out['loss']):.4f}", flush=True)
