"""Per-layer diagnostic of the differentiable reward-FCN path vs torch CPU (train-mode BN)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import creste_public_b200 as cb
from creste_public_b200 import autograd as ag, configs, ops
from creste_public_b200.config import as_cfg
from creste_public_b200.creste.models.blocks.conv import MultiScaleFCN
from oracle import irl_oracle as io
import torch.nn.functional as F
B, H, W = [int(a) for a in sys.argv[1:4]]
case = io.make_case(seed=3, B=B, H=H, W=W)
dev = torch.device("cuda")
port = io.PortMSFCN(); port.load_state_dict(case["state_dict"]); port.train()
cfg = configs.irl_cfg(map_size=(H, W))
net = MultiScaleFCN(as_cfg(cfg["traversability_head"]["net_kwargs"]["reward_cfg"]["net_kwargs"]))
net.load_state_dict(case["state_dict"]); net.train(); net = net.to(dev)
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def rep(name, got, ref):
    got = got.detach().cpu(); ref = nhwc(ref.detach()) if ref.ndim == 4 else ref
    print(f"{name:28s} max|ref|={float(ref.abs().max()):.4g} max|err|={float((got-ref).abs().max()):.3g}")
x = case["input_view"]
with torch.no_grad():
    # layer by layer, feeding the CPU intermediate into the GPU op
    def layer(cl_gpu, cl_cpu, xin, name):
        y_cpu_conv = cl_cpu.conv(xin)
        y_gpu_conv = ag.conv2d(nhwc(xin).to(dev), cl_gpu.conv)
        rep(name + ".conv", y_gpu_conv, y_cpu_conv)
        if hasattr(cl_cpu, "norm"):
            cl_cpu.norm.train()
            y_cpu = torch.relu(cl_cpu.norm(y_cpu_conv))
            y_gpu = ag.batch_norm(nhwc(y_cpu_conv).to(dev), cl_gpu.norm, relu=True)
            rep(name + ".bn_relu", y_gpu, y_cpu)
            return y_cpu
        return torch.relu(y_cpu_conv)
    h = layer(net.prepool[0], port.prepool[0], x, "prepool0")
    h = layer(net.prepool[1], port.prepool[1], h, "prepool1")
    s = layer(net.skip[0], port.skip[0], h, "skip0")
    s = layer(net.skip[1], port.skip[1], s, "skip1")
    t = F.max_pool2d(h, 2, 2)
    rep("maxpool", ag.MaxPool2Fn.apply(nhwc(h).to(dev)), t)
    t = layer(net.trunk[1], port.trunk[1], t, "trunk1")
    t2 = torch.relu(port.trunk[2](t)); rep("trunk2.bn", ag.relu(ag.batch_norm(nhwc(t).to(dev), net.trunk[2])), t2)
    t = layer(net.trunk[4], port.trunk[4], t2, "trunk4")
    t2 = torch.relu(port.trunk[5](t)); rep("trunk5.bn", ag.relu(ag.batch_norm(nhwc(t).to(dev), net.trunk[5])), t2)
    u = F.interpolate(t2, scale_factor=2, mode="bilinear", align_corners=False)
    rep("upsample", ag.Up2Fn.apply(nhwc(t2).to(dev)), u)
    c = torch.cat([u, s], 1)
    r = layer(net.postpool[0], port.postpool[0], c, "postpool0")
port2 = io.PortMSFCN(); port2.load_state_dict(case["state_dict"]); port2.train()
net.load_state_dict(case["state_dict"])
r_cpu = port2(x)
r_gpu = net(x.to(dev).requires_grad_(True))
print("end-to-end r: max|r|", float(r_cpu.abs().max()), "max|err|", float((r_gpu.detach().cpu() - r_cpu.detach()).abs().max()))
