"""Isolate the faulty backward op: CPU torch intermediates + upstream grads -> each GPU Function."""
import os, sys, torch, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch.nn.functional as F
from oracle import irl_oracle as io
from creste_public_b200 import autograd as ag
B, H, W = [int(a) for a in sys.argv[1:4]]
dev = torch.device("cuda")
case = io.make_case(seed=3, B=B, H=H, W=W)
net = io.PortMSFCN(); net.load_state_dict(case["state_dict"]); net.train()
def nhwc(t): return t.detach().permute(0, 2, 3, 1).contiguous().to(dev)
def rep(name, got, ref):
    ref = ref.detach(); got = got.detach().cpu()
    if ref.ndim == 4 and got.shape != ref.shape: ref = ref.permute(0, 2, 3, 1)
    e = (got - ref).abs()
    idx = np.unravel_index(int(e.argmax()), e.shape)
    print(f"{name:34s} max|ref|={float(ref.abs().max()):.4g} max|err|={float(e.max()):.3g} at {tuple(int(i) for i in idx)}")
x0 = case["input_view"]
# forward on CPU keeping intermediates
h0 = net.prepool[0](x0); h1 = net.prepool[1](h0); h1.retain_grad()
s0 = net.skip[0](h1)
mp = F.max_pool2d(h1, 2, 2); mp.retain_grad()
t1 = net.trunk[1](mp); t1.retain_grad()
s0.retain_grad()
rest = torch.relu(net.trunk[2](t1)); rest = torch.relu(net.trunk[5](net.trunk[4](rest)))
up = F.interpolate(rest, scale_factor=2, mode="bilinear", align_corners=False)
r = net.postpool(torch.cat([up, net.skip[1](s0)], 1))
torch.manual_seed(0)
wts = torch.randn_like(r)
(r * wts).sum().backward()
g_h1, g_mp, g_t1, g_s0 = h1.grad, mp.grad, t1.grad, s0.grad
with torch.no_grad():
    # dgrad of trunk1 conv+relu: g at conv output = relu_bwd(g_t1, t1)
    gm = ag.ReluBwdFn.apply(nhwc(g_t1), nhwc(t1))
    w = net.trunk[1].conv.weight.detach().to(dev)
    d_mp = ag.Conv2dFn.apply(gm, ag._flip_t(w), 1, 1)
    rep("dgrad trunk1 -> g(maxpool out)", d_mp, g_mp)
    d_h1_pool = ag.MaxPoolBwdFn.apply(nhwc(g_mp), nhwc(h1))
    # skip0: conv -> bn -> relu; take torch's grad at conv output via autograd on CPU
    with torch.enable_grad():
        cin = h1.detach().clone().requires_grad_(True)
        cout = net.skip[0].conv(cin); cout.retain_grad()
        y = torch.relu(net.skip[0].norm(cout)); (y * g_s0).sum().backward()
    d_h1_skip = ag.Conv2dFn.apply(nhwc(cout.grad), ag._flip_t(net.skip[0].conv.weight.detach().to(dev)), 1, 1)
    rep("dgrad skip0 -> g(h1) skip part", d_h1_skip, cin.grad)
    pin = h1.detach().clone().requires_grad_(True)
    (F.max_pool2d(pin, 2, 2) * g_mp).sum().backward() if False else None
pin = h1.detach().clone().requires_grad_(True)
(F.max_pool2d(pin, 2, 2) * g_mp).sum().backward()
rep("maxpool bwd -> g(h1) pool part", d_h1_pool, pin.grad)
rep("sum", d_h1_pool + d_h1_skip, g_h1)
# BN backward of skip0 on GPU given g_s0
import creste_public_b200 as cb
from creste_public_b200.creste.models.blocks.conv import ConvLayer
