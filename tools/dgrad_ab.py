import os, sys, torch
sys.path.insert(0, "/root/repo")
from creste_public_b200 import ops
torch.manual_seed(0)
for (N, H, W, C, R) in [(4, 128, 240, 144, 3), (4, 64, 120, 240, 5), (2, 33, 61, 40, 5), (3, 16, 30, 1152, 3)]:
    lo = (R - 1) // 2; pad = (lo, R - 1 - lo, lo, R - 1 - lo)
    g = torch.randn(N, H, W, C, device="cuda"); w = torch.randn(R * R, C, device="cuda")
    os.environ.pop("CRESTE_NO_DWDGRAD_TILE", None)
    a = ops.dwconv_dgrad(g, w, (N, H, W, C), R, 1, pad)
    os.environ["CRESTE_NO_DWDGRAD_TILE"] = "1"
    b = ops.dwconv_dgrad(g, w, (N, H, W, C), R, 1, pad)
    os.environ.pop("CRESTE_NO_DWDGRAD_TILE", None)
    print((N, H, W, C, R), "bit-identical" if torch.equal(a, b) else f"DIFF {float((a-b).abs().max())}")
