"""Diagnostic: CUDA forward vs the CPU oracle (and the fp64-exact evaluation) per output key.
Usage: python tools/fwd_check.py H W [profile] [precision]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import creste_public_b200 as cb  # noqa: E402
from oracle import net_oracle, synth  # noqa: E402

H, W = int(sys.argv[1]), int(sys.argv[2])
prof = sys.argv[3] if len(sys.argv) > 3 else "peaky"
prec = sys.argv[4] if len(sys.argv) > 4 else "fp32"
cb.set_precision(prec)
model = cb.build_maxentirl(image_size=(H, W)).eval()
sd = synth.seeded_state_dict(model.state_dict(), 0, prof)
model.load_state_dict(sd)
model = model.cuda()
rgbd, p2p = synth.net_inputs(H, W, 1)
t = time.time()
ref = net_oracle.forward(sd, rgbd, p2p)
ref64 = net_oracle.forward(sd, rgbd, p2p, encoder_fp64=True)
print(f"oracle cpu {time.time() - t:.2f}s")
with torch.no_grad():
    out = model((rgbd.cuda(), p2p.cuda()))
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(3):
        out = model((rgbd.cuda(), p2p.cuda()))
    torch.cuda.synchronize()
    print(f"cuda {(time.time() - t) / 3 * 1e3:.2f} ms/frame")
for k, v in out.items():
    r = ref[k]
    o = v.detach().cpu()
    if o.dtype == torch.int64:
        print(f"{k:34s} mismatches {(o != r.view_as(o)).sum().item()} / {o.numel()}")
    else:
        r = r.view_as(o)
        y = (ref64[k].view_as(o) - r).abs()
        print(f"{k:34s} max|ref| {r.abs().max():9.3e}  max err {(o - r).abs().max():9.3e}  "
              f"mean err {(o - r).abs().mean():9.3e} | yard max {y.max():9.3e} mean {y.mean():9.3e}")
