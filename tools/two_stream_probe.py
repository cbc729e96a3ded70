"""Probe (round 2d; result: NOT adopted, see profiles/r2d_experiments.md): does the eval forward gain from the batch cut
into sub-batches that run on separate streams (the memory- / latency-bound encoder kernels of one sub-batch beside the
tensor-core convs of another), eagerly and as one CUDA graph?  The shipped forward stays single-stream.
CUDA events on the launching stream; no profiler.  Usage: python tools/two_stream_probe.py [precision]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import creste_public_b200 as cb  # noqa: E402
from creste_public_b200 import engine  # noqa: E402
import synth_data as synth  # noqa: E402


_SIDE_STREAMS = {}


def fork_join(fns):
    """Run independent callables on separate CUDA streams (fork from / join into the current stream): branch 0 on
    the current stream, the others on side streams private to it.  Tensors a branch allocates come from its stream's
    allocator pool; every branch starts after the fork event and the caller's stream waits for all of them, so reuse of
    those blocks is ordered.  Capturable in a CUDA graph: fork / join become graph dependencies."""
    if len(fns) < 2:
        return [f() for f in fns]
    main = torch.cuda.current_stream()
    key = (main.device.index, main.cuda_stream, len(fns) - 1)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=main.device) for _ in range(len(fns) - 1)]
    sides = _SIDE_STREAMS[key]
    ev = torch.cuda.Event()
    ev.record(main)
    out = [None] * len(fns)
    for i in range(1, len(fns)):
        sides[i - 1].wait_event(ev)
        with torch.cuda.stream(sides[i - 1]):
            out[i] = fns[i]()
    out[0] = fns[0]()
    for st in sides:
        main.wait_stream(st)
    return out


prec = sys.argv[1] if len(sys.argv) > 1 else "3xfp16"
cb.set_precision(prec)
H, W = 512, 960
B = 8
model = cb.build_maxentirl(image_size=(H, W)).eval()
model.load_state_dict(synth.seeded_state_dict(model.state_dict(), 0, "peaky"))
model = model.cuda()
x = torch.rand(B, 1, 4, H, W, device="cuda")
x[:, :, 3] *= 20000
p2p = torch.from_numpy(synth.make_p2p(H, W)).view(1, 1, 4, 4).repeat(B, 1, 1, 1).cuda()


def timed(fn, n=10):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def whole():
    return model((x, p2p))


def split(parts):
    xs, ps = x.chunk(parts), p2p.chunk(parts)
    return fork_join([(lambda i=i: model((xs[i], ps[i]))) for i in range(parts)])


ms = timed(whole)
print(f"{prec} B={B} one call, one stream: {ms:.2f} ms/step = {B * 1e3 / ms:.1f} frames/s", flush=True)
for parts in (2, 4):
    ms = timed(lambda: split(parts))
    print(f"{prec} B={B} as {parts} sub-batches on {parts} streams: {ms:.2f} ms/step = {B * 1e3 / ms:.1f} frames/s",
          flush=True)

# the same under CUDA-graph replay (no host launch cost: the eager 4-way split is host-bound at ~33 us per launch)
for parts in (1, 2, 4):
    if parts == 1:
        g = engine.GraphedForward(lambda a, b: model((a, b)), (x, p2p))
    else:
        g = engine.GraphedForward(lambda a, b: fork_join(
            [(lambda i=i: model((a.chunk(parts)[i], b.chunk(parts)[i]))) for i in range(parts)]), (x, p2p))
    ms = timed(lambda: g(x, p2p))
    print(f"{prec} B={B} GRAPH, {parts} sub-batch stream(s): {ms:.2f} ms/step = {B * 1e3 / ms:.1f} frames/s", flush=True)
    del g
