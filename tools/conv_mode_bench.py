"""Discriminate MMA-bound vs operand-fill-bound: the same conv in single-pass TF32, 3xTF32, 3xFP16."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops
torch.manual_seed(0)
for (N, C, H, W, K, R) in [(8, 496, 128, 240, 496, 3), (8, 256, 256, 256, 128, 3)]:
    x = torch.randn(N, H, W, C, device="cuda")
    w = torch.randn(K, C, R, R, device="cuda") / (C * R * R) ** 0.5
    pad = (R // 2,) * 4
    for mode in ("tf32", "3xtf32", "3xfp16"):
        wp = ops.pack_conv_weight_f16(w) if mode == "3xfp16" else ops.pack_conv_weight_tc(w, split=(mode == "3xtf32"))
        for _ in range(3):
            ops.conv2d(x, wp, K, R, R, 1, pad, precision=mode)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            ops.conv2d(x, wp, K, R, R, 1, pad, precision=mode)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{mode:7s} C{C}->K{K} {R}x{R} @{H}x{W} B{N}: {ms:.3f} ms (incl. operand pre-pass)  {2*N*H*W*K*C*R*R/ms/1e9:.0f} TFLOP/s")
