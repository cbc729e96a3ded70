"""Forward-only timing at B=8 / B=1 in a given precision + per-op breakdown (CUDA events; no profiler).
Usage: python tools/fwd_profile.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb  # noqa: E402
import synth_data as synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "3xfp16"
cb.set_precision(prec)
H, W = 512, 960
model = cb.build_maxentirl(image_size=(H, W)).eval()
model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
model = model.cuda()
for B in (8, 1):
    x = torch.rand(B, 1, 4, H, W, device="cuda")
    x[:, :, 3] *= 20000
    p2p = torch.from_numpy(synth.make_p2p(H, W)).view(1, 1, 4, 4).repeat(B, 1, 1, 1).cuda()
    with torch.no_grad():
        for _ in range(3):
            model((x, p2p))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            model((x, p2p))
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{prec} B={B}: {ms:.2f} ms/step = {B * 1e3 / ms:.1f} frames/s")
