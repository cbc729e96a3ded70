"""Head-only IRL step micro-benchmark: python tools/irl_bench.py B Hm Wm [precision] [steps]"""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb
from creste_public_b200 import configs, _lib
from creste_public_b200.config import as_cfg
from creste_public_b200.creste.train_traversability import HeadStep
from creste_public_b200.creste.utils.loss_utils import LossManager
from oracle import synth, net_oracle
B, Hm, Wm = [int(a) for a in sys.argv[1:4]]
prec = sys.argv[4] if len(sys.argv) > 4 else "3xtf32"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
cb.set_precision(prec)
dev = torch.device("cuda")
cfg = configs.irl_cfg(image_size=(64, 96), map_size=(Hm, Wm), solve_mdp=True, action_horizon=50)
model = cb.build_maxentirl(cfg).to(dev)
model.backbone.eval(); model.traversability_head.train()
step = HeadStep(model, LossManager(as_cfg(cfg)))
g = torch.Generator(device=dev).manual_seed(0)
feat = {k: torch.randn(B, c, 4 * Hm, 2 * Wm, device=dev, generator=g) for k, c in
        (("inpainting_sam_preds", 32), ("inpainting_sam_dynamic_preds", 6), ("elevation_preds", 2))}
expert = torch.from_numpy(synth.expert_poses(B, 50, 4 * Hm, 2 * Wm, 1)).to(dev)
cfs = synth.counterfactuals(expert.cpu().numpy(), every=2, shift=0.12 * 2 * Wm)
fov = torch.from_numpy(np.ascontiguousarray(net_oracle.trapezoid_fov_mask(4 * Hm, 2 * Wm, 70, 70, 7 * Wm / 128, 200 * Wm / 128)))
fov = fov.unsqueeze(0).repeat(B, 1, 1).to(dev)
for _ in range(2):
    loss, out, meta = step(feat, expert, fov, cfs)
torch.cuda.synchronize()
n0 = _lib.lib().creste_launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(steps):
    loss, out, meta = step(feat, expert, fov, cfs)
b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
ms = a.elapsed_time(b) / steps
print(f"IRL head step B={B} {Hm}x{Wm} {prec}: {ms:.2f} ms/step (wall {(t1 - t0) / steps * 1e3:.2f}), "
      f"{1e3 / ms:.1f} steps/s, loss={float(loss):.5f}, K={int(model.traversability_head.last_vi_info[0])}, "
      f"launches/step={(_lib.lib().creste_launch_count() - n0) / steps:.0f}")
