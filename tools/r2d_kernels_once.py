"""One launch each of the kernels rewritten in round 2d, at the sizes the training steps run them, for
    ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|wgrad_strided_c4|chan_reduce_small|dwconv_fwd_xb|dwconv_dgrad_s2|svf_kernel|chan_axpby_act" \
        -o gpurun_out/r2d_kernels python tools/r2d_kernels_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def rn(*s):
    return torch.randn(*s, device=dev, generator=g)


# weight gradient of the reward FCN's 5x5 40 -> 64 layer (tap groups: 4 taps per CTA share the g tile)
x, gy = rn(8, 256, 256, 40), rn(8, 256, 256, 64)
ops.conv2d_wgrad_tc_presplit(ops.split_f16(x), ops.split_f16(gy), 5, 5, (2, 2, 2, 2))
# stem weight gradient at B = 16
ops.wgrad_strided(rn(16, 512, 960, 4), rn(16, 256, 480, 32), 3, 3, 2, (0, 1, 0, 1))
# BatchNorm backward on a 32-channel layer at 256x480 (small-C reduction) + the fused second pass
x, gy = rn(16, 256, 480, 32), rn(16, 256, 480, 32)
a, b, q, r = rn(32), rn(32), rn(32), rn(32)
ops.bn_act_bwd(gy, x, a, b, "swish", want_gu=False)
ops.chan_axpby_act(gy, x, a, b, "swish", a, q, r, want_amax=True)
# depthwise forward (x-blocked) and stride-2 data gradient
w = rn(9, 96)
ops.dwconv_fwd(rn(16, 256, 480, 96), w, 3, 2, (0, 1, 0, 1))
ops.dwconv_dgrad(rn(16, 128, 240, 96), w, (16, 256, 480, 96), 3, 2, (0, 1, 0, 1))
torch.cuda.synchronize()
