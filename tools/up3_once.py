"""One up3 conv (3x3 496->496 @128x240, B=8, 3xfp16) a few times, in the form the forward runs it since round 2b: the
operand arrives pre-split from the producing conv's epilogue and the output is written as the next conv's operand
(creste_conv2d_presplit_split_out).  The workload of the `ncu --set full` capture behind bench.py's roofline.traffic:
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -o gpurun_out/r2b_up3_conv \
      python tools/up3_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
x = torch.randn(B, 128, 240, 496, device="cuda").relu_()
w = torch.randn(496, 496, 3, 3, device="cuda") / (496 * 9) ** 0.5
wp = ops.pack_conv_weight_f16(w)
bound = (float(w.abs().sum(dim=(1, 2, 3)).max()) * 1.001, 0.0)
xs = ops.conv2d(x, wp, 496, 3, 3, 1, (1, 1, 1, 1), act="relu", precision="3xfp16", split_out=("only",) + bound)  # launch 0
for _ in range(4):
    ops.conv2d_presplit(xs, wp, 496, 3, 3, 1, (1, 1, 1, 1), act="relu", precision="3xfp16", split_out=("only",) + bound)
torch.cuda.synchronize()
print("done")
