"""One launch of each memory-bound helper kernel on its forward shape (B = 8), for `ncu --set full -k regex:...`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from creste_public_b200 import ops  # noqa: E402

B, dev = 8, "cuda"
torch.manual_seed(0)
x = torch.randn(B, 256, 256, 128, device=dev)
w = torch.randn(32, 128, device=dev) / 128 ** 0.5
ops.proj_head(x, w, torch.randn(32, device=dev))
for (Cc, K, H, W, gate, act) in [(16, 96, 256, 480, False, "swish"), (144, 24, 128, 240, True, "none")]:
    xx = torch.randn(B, H, W, Cc, device=dev)
    wp = ops.pack_conv_weight(torch.randn(K, Cc, 1, 1, device=dev) / Cc ** 0.5)
    g = torch.rand(B, Cc, device=dev) if gate else None
    ops.conv2d(xx, wp, K, 1, 1, 1, (0, 0, 0, 0), torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev), g, None, act,
               False, "fp32", amax_out=torch.zeros(1, device=dev))
xs = torch.randn(B, 64, 64, 256, device=dev)
ops.upsample_concat_split(None, xs, (256, 256), 4.0, None, xs.abs().max().reshape(1))
xd = torch.randn(B, 256, 480, 32, device=dev)
ops.dwconv_bn_swish(xd, torch.randn(3, 3, 32, device=dev), torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev), 3, 1,
                    (1, 1, 1, 1))
torch.cuda.synchronize()
