"""A/B check + timing of the memory-bound helper kernels of the forward against their HBM roofline.
For every kernel that has an older form behind an environment switch, the two are compared BIT FOR BIT on the shapes of
the 512x960 forward (B = 8) and timed with CUDA events (median of 5).  Usage: python tools/kernel_bench.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from creste_public_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
PEAK = 6464.3
dev = "cuda"
torch.manual_seed(0)


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def report(name, ms_new, ms_old, nbytes, same):
    gbs = nbytes / ms_new / 1e6
    old = f"{ms_old * 1e3:8.1f} us" if ms_old is not None else "       -   "
    print(f"{name:58s} new {ms_new * 1e3:8.1f} us  old {old}  {gbs:7.0f} GB/s = {gbs / PEAK:4.2f} of HBM  "
          f"{'bit-identical' if same else ('DIFFERENT' if same is not None else '')}", flush=True)


def ab(env, fn):
    os.environ.pop(env, None)
    new = fn()
    t_new = timeit(fn)
    os.environ[env] = "1"
    old = fn()
    t_old = timeit(fn)
    os.environ.pop(env, None)
    return new, old, t_new, t_old


# ---- 1x1 convs of the MBConv blocks (exact fp32)
for (Cc, K, H, W, gate, res, act) in [(32, 16, 256, 480, True, False, "none"), (16, 96, 256, 480, False, False, "swish"),
                                      (96, 24, 128, 240, True, False, "none"), (24, 144, 128, 240, False, False, "swish"),
                                      (144, 24, 128, 240, True, True, "none"), (144, 40, 64, 120, True, False, "none"),
                                      (40, 240, 64, 120, False, False, "swish"), (128, 128, 128, 240, False, False, "relu")]:
    x = torch.randn(B, H, W, Cc, device=dev)
    w = torch.randn(K, Cc, 1, 1, device=dev) / Cc ** 0.5
    wp = ops.pack_conv_weight(w)
    sc, sh = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev)
    g = torch.rand(B, Cc, device=dev) if gate else None
    r = torch.randn(B, H, W, K, device=dev) if res else None
    am = torch.zeros(1, device=dev)
    f = lambda: ops.conv2d(x, wp, K, 1, 1, 1, (0, 0, 0, 0), sc, sh, g, r, act, False, "fp32", amax_out=am)  # noqa: E731
    new, old, tn, to = ab("CRESTE_NO_CONV1X1", f)
    nbytes = 4 * B * H * W * (Cc + K + (K if res else 0))
    report(f"conv1x1 C{Cc}->K{K} @{H}x{W} act={act} gate={int(gate)} res={int(res)}", tn, to, nbytes, torch.equal(new, old))

# ---- bilinear up-sampling + concat (fp32 and split forms)
for (Cs, Cx, Hi, Wi, F, x_first) in [(0, 256, 64, 64, 4, False), (64, 64, 64, 64, 4, False), (112, 320, 16, 30, 2, False),
                                     (40, 192, 32, 60, 2, False), (24, 472, 64, 120, 2, False), (128, 128, 32, 32, 2, True)]:
    x = torch.randn(B, Hi, Wi, Cx, device=dev)
    skip = torch.randn(B, Hi * F, Wi * F, Cs, device=dev) if Cs else None
    f = lambda: ops.upsample_concat(skip, x, (Hi * F, Wi * F), float(F), x_first)  # noqa: E731
    new, old, tn, to = ab("CRESTE_NO_UPSAMPLE_BLOCK", f)
    nbytes = 4 * B * (Hi * Wi * Cx + Hi * F * Wi * F * (2 * Cs + Cx))
    report(f"upsample_concat x{F} Cs{Cs}+Cx{Cx} @{Hi}x{Wi}", tn, to, nbytes, torch.equal(new, old))
    ax = x.abs().max().reshape(1)
    asx = skip.abs().max().reshape(1) if Cs else None
    f2 = lambda: ops.upsample_concat_split(skip, x, (Hi * F, Wi * F), float(F), asx, ax, x_first)  # noqa: E731
    new, old, tn, to = ab("CRESTE_NO_UPSAMPLE_BLOCK", f2)
    same = torch.equal(new.hi.view(torch.int16), old.hi.view(torch.int16)) and \
        torch.equal(new.lo.view(torch.int16), old.lo.view(torch.int16)) and torch.equal(new.scal, old.scal)
    nbytes = 4 * B * (Hi * Wi * Cx + Hi * F * Wi * F * (2 * Cs + Cx))
    report(f"upsample_concat_split x{F} Cs{Cs}+Cx{Cx} @{Hi}x{Wi}", tn, to, nbytes, same)

# ---- layout changes at the module boundary
x = torch.rand(B, 4, 512, 960, device=dev)
y = ops.nchw_to_nhwc(x)
report("nchw_to_nhwc C4 @512x960", timeit(lambda: ops.nchw_to_nhwc(x)), None, 2 * x.numel() * 4,
       torch.equal(y, x.permute(0, 2, 3, 1).contiguous()))
for Cc in (64, 128):
    x = torch.randn(B, 128, 240, Cc, device=dev)
    y = ops.nhwc_to_nchw(x)
    report(f"nhwc_to_nchw C{Cc} @128x240", timeit(lambda: ops.nhwc_to_nchw(x)), None, 2 * x.numel() * 4,
           torch.equal(y, x.permute(0, 3, 1, 2).contiguous()))
    z = ops.nchw_to_nhwc(y)
    report(f"nchw_to_nhwc C{Cc} @128x240", timeit(lambda: ops.nchw_to_nhwc(y)), None, 2 * x.numel() * 4, torch.equal(z, x))

# ---- projection heads (1x1 conv K <= 32 + both NCHW copies)
x = torch.randn(B, 256, 256, 128, device=dev)
for K in (32, 6, 2):
    w = torch.randn(K, 128, device=dev) / 128 ** 0.5
    bias = torch.randn(K, device=dev)
    pred, pred_nchw, x_nchw = ops.proj_head(x, w, bias)
    ref = ops.conv2d(x, ops.pack_conv_weight(w.view(K, 128, 1, 1).contiguous()), K, 1, 1, 1, (0, 0, 0, 0), None, bias)
    same = torch.equal(pred, ref) and torch.equal(pred_nchw, ref.permute(0, 3, 1, 2).contiguous()) and \
        torch.equal(x_nchw, x.permute(0, 3, 1, 2).contiguous())
    report(f"proj_head C128->K{K} @256x256", timeit(lambda: ops.proj_head(x, w, bias)), None,
           4 * B * 256 * 256 * (2 * 128 + 2 * K), same)

# ---- depthwise conv + BN + swish (+ SE partial sums): tiled kernel vs the x-blocked one (CRESTE_NO_DWTILE)
for (Cc, H, W, R, st) in [(32, 256, 480, 3, 1), (96, 256, 480, 3, 2), (144, 128, 240, 3, 1), (144, 128, 240, 5, 2),
                          (240, 64, 120, 5, 1), (240, 64, 120, 3, 2), (480, 32, 60, 3, 1), (672, 32, 60, 5, 1),
                          (672, 32, 60, 5, 2), (1152, 16, 30, 5, 1), (1152, 16, 30, 3, 1)]:
    x = torch.randn(B, H, W, Cc, device=dev)
    w = torch.randn(R * R, Cc, device=dev) / R
    sc, sh = torch.rand(Cc, device=dev) + 0.5, torch.randn(Cc, device=dev) * 0.1
    tot = (R - 1) if st == 1 else max(R - st, 0)
    pad = (tot // 2, tot - tot // 2) * 2
    f = lambda: ops.dwconv_bn_swish(x, w, sc, sh, R, st, pad)  # noqa: E731
    new, old, tn, to = ab("CRESTE_NO_DWTILE", f)
    same = torch.equal(new[0], old[0]) and \
        float((new[1].sum(1) - old[1].sum(1)).abs().max()) <= 1e-5 * float(old[1].sum(1).abs().max())
    P, Q = new[0].shape[1], new[0].shape[2]
    report(f"dwconv C{Cc} @{H}x{W} k{R} s{st}", tn, to, 4 * B * Cc * (H * W + P * Q), same)
