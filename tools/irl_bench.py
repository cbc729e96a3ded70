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
if os.environ.get("PROFILE"):
    import collections, types
    from creste_public_b200 import ops
    rec = []
    def wrap(name, fn):
        def w(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(*a, **k); e1.record()
            tag = name
            if name in ("conv2d", "conv2d_wgrad"):
                tag = f"{name} {tuple(a[0].shape)}->{a[2] if name == 'conv2d' else tuple(a[1].shape)[-1]} k{a[3] if name == 'conv2d' else a[2]}"
            rec.append((tag, name, e0, e1)); return r
        return w
    for n in dir(ops):
        f = getattr(ops, n)
        if isinstance(f, types.FunctionType) and not n.startswith("_") and n not in ("conv_desc", "tc_supported", "tc_layout", "pack_conv_weight", "pack_conv_weight_tc", "rna_tf32", "maxpool2", "upsample2"):
            setattr(ops, n, wrap(n, f))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); step(feat, expert, fov, cfs); t1.record(); torch.cuda.synchronize()
    tot = t0.elapsed_time(t1)
    by = collections.defaultdict(lambda: [0, 0.0]); byt = collections.defaultdict(lambda: [0, 0.0])
    for tag, name, e0, e1 in rec:
        ms = e0.elapsed_time(e1); by[name][0] += 1; by[name][1] += ms; byt[tag][0] += 1; byt[tag][1] += ms
    print(f"profiled step: {tot:.2f} ms, {len(rec)} op calls, sum of ops {sum(v[1] for v in by.values()):.2f} ms")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:14]: print(f"  {k:22s} n={v[0]:4d} {v[1]:8.2f} ms")
    for k, v in sorted(byt.items(), key=lambda kv: -kv[1][1])[:12]: print(f"    {k:60s} n={v[0]:3d} {v[1]:7.2f} ms")
    sys.exit(0)
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
