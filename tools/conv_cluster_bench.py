"""up3-like conv timing for different CRESTE_TC_CLUSTER values (set in the environment)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops
import torch.nn.functional as F
torch.manual_seed(0)
for (N, C, H, W, K, R) in [(8, 496, 128, 240, 496, 3), (8, 256, 256, 256, 128, 3), (8, 320, 128, 128, 256, 3), (8, 64, 128, 128, 64, 3)]:
    x = torch.randn(N, H, W, C, device="cuda")
    w = torch.randn(K, C, R, R, device="cuda") / (C * R * R) ** 0.5
    wp = ops.pack_conv_weight_f16(w)
    pad = (R // 2,) * 4
    for _ in range(3):
        out = ops.conv2d(x, wp, K, R, R, 1, pad, precision="3xfp16")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        out = ops.conv2d(x, wp, K, R, R, 1, pad, precision="3xfp16")
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    # correctness on a slice
    ref = F.conv2d(x[:1].permute(0, 3, 1, 2), w, padding=R // 2).permute(0, 2, 3, 1)
    err = float((out[:1] - ref).abs().max() / ref.abs().max())
    print(f"cluster={os.environ.get('CRESTE_TC_CLUSTER','4')} C{C}->K{K} {R}x{R} @{H}x{W} B{N}: {ms:.3f} ms  {2*N*H*W*K*C*R*R/ms/1e9:.0f} TFLOP/s  relerr {err:.2e}")
