import os, sys, torch, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from oracle import irl_oracle as io
from creste_public_b200 import configs
from creste_public_b200.config import as_cfg
from creste_public_b200.creste.models.blocks.conv import MultiScaleFCN
from creste_public_b200.creste.utils.loss_utils import MaxEntIRLLoss
B, H, W = [int(a) for a in sys.argv[1:4]]
dev = torch.device("cuda")
case = io.make_case(seed=3, B=B, H=H, W=W)
for rw, mw in ((0.01, 1.0), (0.0, 1.0), (1.0, 0.0)):
    net = io.PortMSFCN().double(); net.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in case["state_dict"].items()}); net.train()
    c64 = dict(case); c64["input_view"] = case["input_view"].double(); c64["exp_svf"] = case["exp_svf"].double()
    p64 = io.run_steps(net, io.PortLoss(case["map_size"], reward_weight=rw, maxent_weight=mw), c64, 1)
    cfg = configs.irl_cfg(map_size=(H, W))
    cfg["loss"][0]["reward_weight"] = rw; cfg["loss"][0]["maxent_weight"] = mw
    n = MultiScaleFCN(as_cfg(cfg["traversability_head"]["net_kwargs"]["reward_cfg"]["net_kwargs"]))
    n.load_state_dict(case["state_dict"]); n.train(); n = n.to(dev)
    ours = io.run_steps(n, MaxEntIRLLoss(as_cfg(cfg["loss"][0])), case, 1, device=dev)
    print(f"--- reward_weight={rw} maxent_weight={mw}: loss ours {ours['loss'][0]:.6g} f64 {p64['loss'][0]:.6g} pen {ours['reward_penalty'][0]:.6g} {p64['reward_penalty'][0]:.6g}")
    for k in p64["grads"]:
        g = p64["grads"][k]
        print(f"  {k:28s} max|g|={np.abs(g).max():.3g} relerr={np.abs(ours['grads'][k]-g).max()/max(np.abs(g).max(),1e-12):.3g}")
