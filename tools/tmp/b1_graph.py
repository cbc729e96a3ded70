import os, sys, torch
sys.path.insert(0, ".")
import creste_public_b200 as cb
from creste_public_b200 import engine
import synth_data as synth
cb.set_precision("3xfp16")
H, W = 512, 960
model = cb.build_maxentirl(image_size=(H, W)).eval()
model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
model = model.cuda()
for B in (1, 2):
    x = torch.rand(B, 1, 4, H, W, device="cuda"); x[:, :, 3] *= 20000
    p2p = torch.from_numpy(synth.make_p2p(H, W)).view(1, 1, 4, 4).repeat(B, 1, 1, 1).cuda()
    g = engine.GraphedForward(lambda a, b: model((a, b)), (x, p2p))
    for _ in range(5): g(x, p2p)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): g(x, p2p)
    b.record(); torch.cuda.synchronize()
    print(f"B={B} graph: {a.elapsed_time(b) / 20:.3f} ms")
