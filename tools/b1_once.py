"""A few B = 1 eval forwards at 512x960 (3xfp16) for a launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file out.csv python tools/b1_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb  # noqa: E402
import synth_data as synth  # noqa: E402

cb.set_precision(sys.argv[1] if len(sys.argv) > 1 else "3xfp16")
H, W = 512, 960
model = cb.build_maxentirl(image_size=(H, W)).eval()
model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
model = model.cuda()
x = torch.rand(1, 1, 4, H, W, device="cuda")
x[:, :, 3] *= 20000
p2p = torch.from_numpy(synth.make_p2p(H, W)).view(1, 1, 4, 4).cuda()
with torch.no_grad():
    for _ in range(5):
        model((x, p2p))
torch.cuda.synchronize()
