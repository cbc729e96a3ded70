"""Stage-1 (distillation) training step micro-benchmark: python tools/distill_bench.py B H W [precision] [steps]
PROFILE=1 prints the per-op time breakdown of one step (CUDA events around every ops.* call)."""
import collections, os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb
from creste_public_b200 import configs, _lib, ops
from creste_public_b200.creste.train_pefree import DistillationModel
import synth_data

B, H, W = [int(a) for a in sys.argv[1:4]]
prec = sys.argv[4] if len(sys.argv) > 4 else "3xfp16"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
cb.set_precision(prec)
dev = torch.device("cuda")
torch.manual_seed(0)
m = DistillationModel(configs.distill_cfg((H, W))).to(dev).train()
inputs = {k: v.to(dev) for k, v in synth_data.distill_batch(B, H, W, seed=0).items()}
for _ in range(2):
    out = m.training_step(inputs)
torch.cuda.synchronize()
print(f"warm: loss={float(out['loss']):.4f} peak_mem={torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
if os.environ.get("PROFILE"):
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
    skip = ("conv_desc", "tc_supported", "tc_layout", "pack_conv_weight", "pack_conv_weight_tc", "pack_conv_weight_f16",
            "rna_tf32", "maxpool2", "upsample2", "ptr", "lib", "stream", "check", "wgrad_tc_supported")
    for n in dir(ops):
        f = getattr(ops, n)
        if isinstance(f, types.FunctionType) and not n.startswith("_") and n not in skip:
            setattr(ops, n, wrap(n, f))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); m.training_step(inputs); t1.record(); torch.cuda.synchronize()
    tot = t0.elapsed_time(t1)
    by = collections.defaultdict(lambda: [0, 0.0]); byt = collections.defaultdict(lambda: [0, 0.0])
    for tag, name, e0, e1 in rec:
        ms = e0.elapsed_time(e1); by[name][0] += 1; by[name][1] += ms; byt[tag][0] += 1; byt[tag][1] += ms
    print(f"profiled step: {tot:.2f} ms, {len(rec)} op calls, sum of ops {sum(v[1] for v in by.values()):.2f} ms")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:16]: print(f"  {k:22s} n={v[0]:4d} {v[1]:8.2f} ms")
    for k, v in [kv for kv in sorted(byt.items(), key=lambda kv: -kv[1][1]) if " " in kv[0]][:int(os.environ.get("PROFILE_TAGS", "14"))]: print(f"    {k:64s} n={v[0]:3d} {v[1]:7.2f} ms")
    sys.exit(0)
n0 = _lib.lib().creste_launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(steps):
    out = m.training_step(inputs)
b.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / steps
ms = a.elapsed_time(b) / steps
n1 = _lib.lib().creste_launch_count()
print(f"stage-1 step B={B} {H}x{W} {prec}: {ms:.1f} ms/step (wall {wall:.1f}), {B / ms * 1e3:.1f} frames/s, "
      f"{3 * 657.5 * B * (H * W) / (512 * 960) / ms:.1f} TFLOP/s (3x fwd flop), loss={float(out['loss']):.4f}, "
      f"launches/step={(n1 - n0) / steps:.0f}, peak_mem={torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
