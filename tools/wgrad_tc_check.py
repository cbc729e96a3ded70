"""tcgen05 wgrad (MN-major operands) against torch CPU float64: python tools/wgrad_tc_check.py [swap]"""
import os, sys, time
if len(sys.argv) > 1:
    os.environ["CRESTE_WGRAD_SWAP"] = sys.argv[1]
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops
dev = torch.device("cuda")
g_ = torch.Generator().manual_seed(0)
shapes = [(2, 16, 24, 64, 64, 1), (2, 16, 24, 128, 128, 1), (2, 16, 24, 128, 256, 3), (2, 16, 30, 496, 496, 3),
          (2, 20, 28, 112, 72, 3), (3, 17, 23, 72, 200, 1), (1, 32, 60, 432, 432, 3), (2, 16, 24, 1152, 192, 1), (4, 128, 240, 496, 496, 3)]
for (N, H, W, C, K, R) in shapes:
    pad = (R // 2,) * 4
    x = torch.randn(N, H, W, C, generator=g_)
    gy = torch.randn(N, H, W, K, generator=g_) * 1e-3
    t0 = time.perf_counter()
    dw = ops.conv2d_wgrad_tc(x.to(dev), gy.to(dev), R, R, pad)
    torch.cuda.synchronize()
    big = N * H * W > 20000
    if big:      # too slow for a float64 CPU reference: check against the CUDA-core kernel on a 64 x 64 tile
        ref = ops.conv2d_wgrad(x[..., :64].contiguous().to(dev), gy[..., :64].contiguous().to(dev), R, R, pad).cpu().double()
        got = dw[:64, :64].cpu().double()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        xd, gd = x.to(dev), gy.to(dev)
        e0.record()
        for _ in range(5):
            ops.conv2d_wgrad_tc(xd, gd, R, R, pad)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        extra = f"  {ms:.3f} ms  {2 * N * H * W * C * K * R * R / ms / 1e9:.1f} TFLOP/s (incl. operand split)"
    else:
        w = torch.zeros(K, C, R, R, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(x.permute(0, 3, 1, 2).double(), w, padding=R // 2)
        (ref,) = torch.autograd.grad(y, w, gy.permute(0, 3, 1, 2).double())
        got = dw.cpu().double()
        extra = ""
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"N{N} {H}x{W} C{C}->K{K} k{R}: err/max = {err:.3e}{extra}", flush=True)
