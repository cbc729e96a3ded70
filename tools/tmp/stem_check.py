import sys, torch
sys.path.insert(0, ".")
from creste_public_b200 import ops
torch.manual_seed(0)
x = torch.randn(16, 512, 960, 4, device="cuda"); g = torch.randn(16, 256, 480, 32, device="cuda") * 0.01 + 0.003
pad = (0, 1, 0, 1)
dw = ops.wgrad_strided(x, g, 3, 3, 2, pad)
xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2).double(), (0, 1, 0, 1))
ref = torch.nn.grad.conv2d_weight(xp, (32, 4, 3, 3), g.permute(0, 3, 1, 2).double().contiguous(), stride=2)
print("max rel err vs fp64:", float((dw.double() - ref).abs().max() / ref.abs().max()))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): ops.wgrad_strided(x, g, 3, 3, 2, pad)
a.record()
for _ in range(10): ops.wgrad_strided(x, g, 3, 3, 2, pad)
b.record(); torch.cuda.synchronize()
print("ms per call:", a.elapsed_time(b) / 10, " -> GB/s", (x.numel() + g.numel()) * 4 / (a.elapsed_time(b) / 10 * 1e-3) / 1e9)
