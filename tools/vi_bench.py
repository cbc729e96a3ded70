"""VI micro-benchmark: python tools/vi_bench.py B H W"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from creste_public_b200 import ops
from oracle import synth
B, H, W = [int(a) for a in sys.argv[1:4]]
r = torch.from_numpy(synth.vi_inputs(7, B, H, W)).cuda()
for _ in range(2):
    v, q, pi, info = ops.vi_solve(r, 0.99, 1e-3)
torch.cuda.synchronize()
os.environ.pop("CRESTE_VI_DEBUG", None)
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); v, q, pi, info = ops.vi_solve(r, 0.99, 1e-3); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
K = int(info[0]); ms = sorted(ts)[len(ts)//2]
gb = B*H*W*(12.0*K+76)/1e9
print(f"B={B} {H}x{W} K={K} {ms:.3f} ms  {ms*1e3/K:.2f} us/sweep  {gb/(ms/1e3):.0f} GB/s-equivalent ({gb/(ms/1e3)/6536*100:.0f}% of HBM copy peak)")
