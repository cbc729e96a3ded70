"""Compare retained intermediate gradients of the real GPU autograd pipeline vs the CPU port."""
import os, sys, torch, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch.nn.functional as F
from oracle import irl_oracle as io
from creste_public_b200 import autograd as ag, configs
from creste_public_b200.config import as_cfg
from creste_public_b200.creste.models.blocks.conv import MultiScaleFCN
B, H, W = [int(a) for a in sys.argv[1:4]]
dev = torch.device("cuda")
case = io.make_case(seed=3, B=B, H=H, W=W)
port = io.PortMSFCN(); port.load_state_dict(case["state_dict"]); port.train()
cfg = configs.irl_cfg(map_size=(H, W))
net = MultiScaleFCN(as_cfg(cfg["traversability_head"]["net_kwargs"]["reward_cfg"]["net_kwargs"]))
net.load_state_dict(case["state_dict"]); net.train(); net = net.to(dev)
x0 = case["input_view"]
torch.manual_seed(0)
wts = torch.randn(B, 1, H, W)
cpu, gpu = {}, {}
def keep(d, name, t):
    t.retain_grad(); d[name] = t; return t
def cpu_layer(cl, x, name):
    y = keep(cpu, name + ".conv", cl.conv(x))
    if hasattr(cl, "norm"):
        y = cl.norm(y)
    return keep(cpu, name + ".out", torch.relu(y))
xc = x0.clone().requires_grad_(True)
h0 = cpu_layer(port.prepool[0], xc, "prepool0"); h1 = cpu_layer(port.prepool[1], h0, "prepool1")
s0 = cpu_layer(port.skip[0], h1, "skip0"); s1 = cpu_layer(port.skip[1], s0, "skip1")
mp = keep(cpu, "maxpool", F.max_pool2d(h1, 2, 2))
t1 = cpu_layer(port.trunk[1], mp, "trunk1")
t2 = keep(cpu, "trunk2.out", torch.relu(port.trunk[2](t1)))
t4 = cpu_layer(port.trunk[4], t2, "trunk4")
t5 = keep(cpu, "trunk5.out", torch.relu(port.trunk[5](t4)))
up = keep(cpu, "up", F.interpolate(t5, scale_factor=2, mode="bilinear", align_corners=False))
r = cpu_layer(port.postpool[0], torch.cat([up, s1], 1), "postpool0")
(r * wts).sum().backward()
def gpu_layer(cl, x, name):
    y = keep(gpu, name + ".conv", ag.conv2d(x, cl.conv))
    if hasattr(cl, "norm"):
        return keep(gpu, name + ".out", ag.batch_norm(y, cl.norm, relu=True))
    return keep(gpu, name + ".out", ag.relu(y))
xg = x0.to(dev).requires_grad_(True)
g0 = gpu_layer(net.prepool[0], ag.ToNHWC.apply(xg), "prepool0"); g1 = gpu_layer(net.prepool[1], g0, "prepool1")
gs0 = gpu_layer(net.skip[0], g1, "skip0"); gs1 = gpu_layer(net.skip[1], gs0, "skip1")
gmp = keep(gpu, "maxpool", ag.MaxPool2Fn.apply(g1))
gt1 = gpu_layer(net.trunk[1], gmp, "trunk1")
gt2 = keep(gpu, "trunk2.out", ag.relu(ag.batch_norm(gt1, net.trunk[2])))
gt4 = gpu_layer(net.trunk[4], gt2, "trunk4")
gt5 = keep(gpu, "trunk5.out", ag.relu(ag.batch_norm(gt4, net.trunk[5])))
gup = keep(gpu, "up", ag.Up2Fn.apply(gt5))
gr = gpu_layer(net.postpool[0], torch.cat([gup, gs1], -1), "postpool0")
(gr * wts.permute(0, 2, 3, 1).contiguous().to(dev)).sum().backward()
print(f"{'tensor':18s} {'max|v|':>9s} {'verr':>9s} {'max|g|':>9s} {'gerr':>9s} {'#bad':>6s}")
for k in cpu:
    v, g = cpu[k].detach().permute(0, 2, 3, 1), cpu[k].grad.permute(0, 2, 3, 1)
    vg, gg = gpu[k].detach().cpu(), gpu[k].grad.cpu()
    ge = (gg - g).abs()
    print(f"{k:18s} {float(v.abs().max()):9.3g} {float((vg - v).abs().max()):9.3g} {float(g.abs().max()):9.3g} "
          f"{float(ge.max()):9.3g} {int((ge > 1e-3 * g.abs().max()).sum()):6d}")
print("input grad err", float((xg.grad.cpu() - xc.grad).abs().max()), float(xc.grad.abs().max()))
for (n1, p1), (n2, p2) in zip(port.named_parameters(), net.named_parameters()):
    print(f"  {n1:28s} max|g|={float(p1.grad.abs().max()):.3g} relerr={float((p2.grad.cpu()-p1.grad).abs().max()/p1.grad.abs().max()):.3g}")
