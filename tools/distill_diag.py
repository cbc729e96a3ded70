"""Diagnostics of the stage-1 step parity: per-mode gradient error table against the float64 port."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import distill_oracle as do
from test_distill_cpu import ours_step
from creste_public_b200 import engine, ops
import torch.nn.functional as F

dev = torch.device("cuda")
# 1. conv kernels on the dgrad shapes, gradient-sized magnitudes
g = torch.Generator().manual_seed(1)
for (C, K, R, mag) in [(256, 496, 1, 1e-6), (128, 256, 3, 1e-6), (496, 496, 3, 1e-6), (256, 496, 1, 1.0), (128, 256, 3, 1.0)]:
    x = torch.randn(2, 16, 24, C, generator=g) * mag
    w = torch.randn(K, C, R, R, generator=g) * 0.05
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=R // 2).permute(0, 2, 3, 1)
    for mode in ("fp32", "3xfp16"):
        m = engine.pick_mode(tuple(x.shape), K, R, R, 1, (R // 2,) * 4, mode)
        packed = ops.pack_conv_weight(w.to(dev)) if m == "fp32" else ops.pack_conv_weight_f16(w.to(dev))
        y = ops.conv2d(x.to(dev), packed, K, R, R, 1, (R // 2,) * 4, precision=m)
        print(f"conv C{C}->K{K} k{R} mag {mag:g} {mode}->{m}: err/max = {(y.cpu().double() - ref).abs().max() / ref.abs().max():.2e}")

case = do.make_case()
port = do.port_step(case)
truth = do.port_grads_fp64(case)[0]
res = {}
for mode in ("fp32", "3xfp16", "fp32", "3xfp16"):
    engine.set_precision(mode)
    ours = ours_step(case, device=dev)
    rows = []
    for k, g0 in port["grads"].items():
        t = truth[k]
        err = np.abs(ours["grads"][k] - t).max(); yard = np.abs(g0 - t).max()
        lim = 3 * yard + 5e-4 * np.abs(t).max() + 1e-6
        rows.append((err / lim, k, err, yard, np.abs(t).max()))
    rows.sort(reverse=True)
    print(f"== mode {mode}: loss {ours['loss']:.6f} (port {port['loss']:.6f}); {sum(r[0] > 1 for r in rows)} tensors over the limit")
    for r in rows[:8]:
        print("   ratio %.2f  %-70s err %.3e yard %.3e max %.3e" % r)
    res.setdefault(mode, []).append(ours)
for mode in res:
    a, b = res[mode]
    print(mode, "run-to-run max grad diff:", max(np.abs(a["grads"][k] - b["grads"][k]).max() for k in a["grads"]))
