"""Per-op GPU time of one forward (CUDA events around every ops.* call; no profiler needed).
Usage: python tools/op_profile.py [batch] [precision]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import creste_public_b200 as cb  # noqa: E402
from creste_public_b200 import ops  # noqa: E402
from oracle import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "3xtf32"
cb.set_precision(prec)
H, W = 512, 960
model = cb.build_maxentirl(image_size=(H, W)).eval()
model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
model = model.cuda()
x = torch.rand(B, 1, 4, H, W, device="cuda")
x[:, :, 3] *= 20000
p2p = torch.from_numpy(synth.make_p2p(H, W)).view(1, 1, 4, 4).repeat(B, 1, 1, 1).cuda()
with torch.no_grad():
    for _ in range(2):
        model((x, p2p))
torch.cuda.synchronize()

records = []


def wrap(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        tag = name
        if name == "conv2d":
            xs = a[0].shape
            tag = f"conv2d[{k.get('precision', a[14] if len(a) > 14 else '?')}] C{xs[3]}->K{a[2]} {a[3]}x{a[4]} s{a[5]} @{xs[1]}x{xs[2]}"
        records.append((tag, name, e0, e1))
        return r
    return w


for n in ["conv2d", "dwconv_bn_swish", "se_gate", "upsample_concat", "maxpool2_concat", "nchw_to_nhwc",
          "nhwc_to_nchw", "splat_soft", "frustum_to_bev", "zmlp_concat", "depth_expectation"]:
    setattr(ops, n, wrap(n, getattr(ops, n)))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    t0.record()
    model((x, p2p))
    t1.record()
torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
by_name = collections.defaultdict(float)
by_tag = collections.defaultdict(lambda: [0, 0.0])
for tag, name, e0, e1 in records:
    ms = e0.elapsed_time(e1)
    by_name[name] += ms
    by_tag[tag][0] += 1
    by_tag[tag][1] += ms
print(f"batch {B} precision {prec}: {tot:.2f} ms/step = {tot / B:.2f} ms/frame, {len(records)} op calls")
for k, v in sorted(by_name.items(), key=lambda kv: -kv[1]):
    print(f"  {k:20s} {v:8.2f} ms {100 * v / tot:5.1f}%")
print("top conv shapes:")
for k, v in sorted(by_tag.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {k:60s} n={v[0]:3d} {v[1]:7.2f} ms")
