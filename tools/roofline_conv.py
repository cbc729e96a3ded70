"""The bench.py roofline kernel alone (effnet up3.conv.3: 3x3 496->496 @128x240, B frames), for an
`ncu --set full -k regex:conv_tc_kernel` capture of its DRAM traffic.  python tools/roofline_conv.py [B] [precision]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb
from oracle import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "3xfp16"
cb.set_precision(prec)
model = cb.build_maxentirl(image_size=(512, 960)).eval()
model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
model = model.cuda()
up3 = model.backbone.depthcomp.depthcomp.vision_backbone.model.up3
x = torch.randn(B, 128, 240, 496, device="cuda")
with torch.no_grad():
    for _ in range(4):
        up3._f1(x, act="relu")
torch.cuda.synchronize()
