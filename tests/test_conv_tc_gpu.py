"""GPU parity of the tcgen05 conv path (TMA + UMMA + TMEM) against torch CPU fp32 convolutions.

3xTF32 (precision 1) is the high-precision tensor-core mode: products carry ~21 mantissa bits;
the residual error is the tensor core's round-toward-zero fp32 accumulation, ~K/8 * 2^-25
relative (measured 1e-5 at K = 4464), so the tolerance is 3e-5 * max|ref| (FFMA path: 2e-5).
Single-pass TF32 (precision 2) is a fast mode: its error is measured and bounded loosely (1e-3
relative), never presented as fp32 parity."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # N, C, H, W, K, R, act, gate, residual
    (1, 32, 8, 16, 32, 1, "none", False, False),        # one tile, one k-block
    (1, 64, 16, 16, 64, 1, "relu", False, False),       # 2 tiles, 2 k-blocks
    (1, 64, 16, 24, 96, 3, "relu", False, False),       # 3x3 halo via TMA OOB fill
    (1, 40, 13, 21, 72, 3, "none", False, True),        # ragged tiles, C % 32 != 0, residual
    (2, 256, 12, 20, 128, 3, "relu", False, False),     # depth-head-like
    (1, 288, 16, 30, 96, 1, "relu", False, False),      # fusion conv
    (1, 496, 16, 30, 496, 3, "relu", False, False),     # up3-like: two N tiles of 256
    (1, 432, 16, 15, 432, 3, "relu", False, False),     # N tile 224
    (2, 96, 16, 16, 24, 1, "none", True, True),         # gate + residual (MBConv project)
    (1, 128, 32, 32, 32, 1, "none", False, False),      # proj head K=32
    (1, 320, 16, 16, 256, 3, "swish", False, False),
]


def _ref(case, g):
    N, C, H, W, K, R, act, use_gate, use_res = case
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(K, C, R, R, generator=g) / (C * R * R) ** 0.5
    scale = torch.rand(K, generator=g) + 0.5
    shift = torch.randn(K, generator=g) * 0.1
    gate = torch.rand(N, C, generator=g) if use_gate else None
    xin = x * gate.view(N, C, 1, 1) if use_gate else x
    ref = F.conv2d(xin, w, padding=R // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = torch.randn(ref.shape, generator=g) if use_res else None
    if use_res:
        ref = ref + res
    ref = {"none": lambda t: t, "relu": F.relu, "swish": lambda t: t * torch.sigmoid(t)}[act](ref)
    return x, w, scale, shift, gate, res, ref


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mode,tol", [("3xtf32", 3e-5), ("tf32", 2e-3), ("3xfp16", 3e-5)])
def test_conv_tc_matches_torch(cuda, case, mode, tol):
    from creste_public_b200 import ops
    N, C, H, W, K, R, act, use_gate, use_res = case
    g = torch.Generator().manual_seed(sum(case[:6]))
    x, w, scale, shift, gate, res, ref = _ref(case, g)
    pad = (R // 2,) * 4
    assert ops.tc_supported((N, H, W, C), K, R, R, 1, pad, mode)
    wp = ops.pack_conv_weight_f16(w.to(cuda)) if mode == "3xfp16" else \
        ops.pack_conv_weight_tc(w.to(cuda), split=(mode == "3xtf32"))
    out = ops.conv2d(x.to(cuda).permute(0, 2, 3, 1).contiguous(), wp, K, R, R, 1, pad, scale.to(cuda),
                     shift.to(cuda), gate.to(cuda) if use_gate else None,
                     res.to(cuda).permute(0, 2, 3, 1).contiguous() if use_res else None, act,
                     precision=mode)
    torch.cuda.synchronize()
    out = out.cpu().permute(0, 3, 1, 2)
    err = float((out - ref).abs().max())
    assert err <= tol * float(ref.abs().max()), f"err {err:.3e} vs max {float(ref.abs().max()):.3e}"


@pytest.mark.parametrize("mode", ["3xtf32", "3xfp16"])
@pytest.mark.parametrize("case", [
    # N, C, H, W, K, R, pad (t, b, l, r)
    (2, 96, 32, 48, 64, 7, (3, 3, 3, 3)),       # BEV stem: conv7x7 s2 p3 (inpainting.py:80-90)
    (1, 64, 33, 41, 32, 3, (0, 1, 0, 1)),       # TF-"SAME" static padding, odd size
    (1, 128, 32, 32, 128, 1, (0, 0, 0, 0)),     # 1x1 s2 (ResNet downsample)
])
def test_conv_tc_stride2(cuda, case, mode):
    """stride-2 convs on the tensor-core kernel: TMA element strides pick every 2nd pixel of the box"""
    from creste_public_b200 import ops
    N, C, H, W, K, R, pad = case
    g = torch.Generator().manual_seed(R + C)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(K, C, R, R, generator=g) / (C * R * R) ** 0.5
    ref = F.conv2d(F.pad(x, (pad[2], pad[3], pad[0], pad[1])), w, stride=2)
    assert ops.tc_supported((N, H, W, C), K, R, R, 2, pad, mode)
    wp = ops.pack_conv_weight_f16(w.to(cuda)) if mode == "3xfp16" else ops.pack_conv_weight_tc(w.to(cuda), split=True)
    out = ops.conv2d(x.to(cuda).permute(0, 2, 3, 1).contiguous(), wp, K, R, R, 2, pad, precision=mode)
    out = out.cpu().permute(0, 3, 1, 2)
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 3e-5 * float(ref.abs().max())


@pytest.mark.parametrize("xs,ws", [(1e-4, 1.0), (3e3, 1e-3), (1.0, 50.0)])
def test_conv_3xfp16_dynamic_range(cuda, xs, ws):
    """3xFP16 (precision 4): fp16 hi/lo operands under power-of-two scales (per tensor for the
    activations, per output channel for the weights).  Inputs far outside fp16's range, a wide
    spread inside one tensor and per-channel weight magnitudes over 6 decades must still give
    the 3xTF32 accuracy (both carry 11-bit significands)."""
    from creste_public_b200 import ops
    g = torch.Generator().manual_seed(7)
    N, C, H, W, K, R = 1, 128, 16, 24, 64, 3
    x = torch.randn(N, C, H, W, generator=g) * xs
    x[:, ::7] *= 1e-3                                   # small-magnitude channels next to large ones
    w = torch.randn(K, C, R, R, generator=g) * ws / (C * R * R) ** 0.5
    w *= torch.logspace(-3, 3, K).view(K, 1, 1, 1)      # per-output-channel scale over 6 decades
    ref = F.conv2d(x.double(), w.double(), padding=1)
    out = ops.conv2d(x.to(cuda).permute(0, 2, 3, 1).contiguous(), ops.pack_conv_weight_f16(w.to(cuda)), K, R, R,
                     1, (1, 1, 1, 1), precision="3xfp16").cpu().permute(0, 3, 1, 2).double()
    # per output channel (each has its own magnitude)
    err = (out - ref).abs().amax(dim=(0, 2, 3)) / ref.abs().amax(dim=(0, 2, 3))
    assert float(err.max()) <= 3e-5, err


def test_conv_3xfp16_zero_input(cuda):
    from creste_public_b200 import ops
    x = torch.zeros(1, 16, 16, 64, device=cuda)
    w = torch.randn(32, 64, 1, 1, device=cuda)
    out = ops.conv2d(x, ops.pack_conv_weight_f16(w), 32, 1, 1, 1, (0, 0, 0, 0), precision="3xfp16")
    assert float(out.abs().max()) == 0.0


def test_tc_unsupported_shapes_are_refused(cuda):
    from creste_public_b200 import ops
    assert not ops.tc_supported((1, 32, 32, 4), 32, 3, 3, 2, (0, 1, 0, 1), "3xfp16")   # C = 4 stem
    assert not ops.tc_supported((1, 32, 32, 128), 6, 1, 1, 1, (0, 0, 0, 0), "3xtf32")   # K = 6 head
    x = torch.randn(1, 16, 16, 64, device=cuda)
    w = ops.pack_conv_weight(torch.randn(6, 64, 1, 1, device=cuda))
    with pytest.raises(RuntimeError, match="precision mode"):
        ops.conv2d(x, w, 6, 1, 1, 1, (0, 0, 0, 0), precision="3xtf32")


@pytest.mark.parametrize("C,K1,K2,H,W,R", [(64, 64, 32, 32, 48, 3), (496, 496, 256, 16, 30, 3), (320, 256, 256, 24, 24, 3),
                                           (256, 128, 128, 20, 28, 1)])
def test_split_output_epilogue_feeds_the_next_conv(cuda, C, K1, K2, H, W, R):
    """conv -> BN -> ReLU -> conv with the first conv's epilogue writing the second conv's 3xFP16 operand
    (creste_conv2d_split_out / creste_conv2d_presplit_split_out): the hi / lo halves reconstruct the fp32 output to
    2^-21 of its bound, the fp32 output of the "both" form is untouched, and the chain gives the same result as the
    split pre-pass (the two power-of-two scales differ, so elements that are fp16-subnormal under the looser one may
    differ in their last bits: 1e-6 of the maximum)."""
    import creste_public_b200 as cb
    from creste_public_b200 import ops
    from creste_public_b200.engine import FusedConv
    from torch import nn
    torch.manual_seed(C + K1)
    cb.set_precision("3xfp16")
    try:
        c1, b1 = nn.Conv2d(C, K1, R, padding=R // 2, bias=False).to(cuda), nn.BatchNorm2d(K1).to(cuda).eval()
        c2, b2 = nn.Conv2d(K1, K2, 3, padding=1, bias=False).to(cuda), nn.BatchNorm2d(K2).to(cuda).eval()
        for b in (b1, b2):
            b.running_mean.normal_(0, 0.1); b.running_var.uniform_(0.5, 1.5); b.weight.data.uniform_(0.5, 1.5)
            b.bias.data.normal_(0, 0.1)
        f1, f2 = FusedConv(c1, b1), FusedConv(c2, b2)
        x = torch.randn(2, H, W, C, device=cuda)
        with torch.no_grad():
            mid = f1(x, act="relu")
            ref = f2(mid, act="relu")
            sp = f1(x, act="relu", split_out="only")
            assert isinstance(sp, ops.SplitAct) and f2.split_ok(mid.shape)
            s, inv = float(sp.scal[0]), float(sp.scal[1])
            assert s * inv == 1.0 and float(mid.abs().max()) * s <= 32768.0
            rec = (sp.hi.float() + sp.lo.float() / 2048.0) * inv
            assert float((rec - mid).abs().max()) <= 2.0 ** -21 * (32768.0 * inv)
            out_only = f2(sp, act="relu")
            both = f1(x, act="relu", split_out="both")
            assert torch.equal(both, mid) and isinstance(both._split, ops.SplitAct)
            assert torch.equal(both._split.hi, sp.hi) and torch.equal(both._split.lo, sp.lo)
            out_both = f2(both, act="relu")
            # a SplitAct input and a split output in the same launch (creste_conv2d_presplit_split_out)
            sp2 = f2(sp, act="relu", split_out="only")
            rec2 = (sp2.hi.float() + sp2.lo.float() / 2048.0) * float(sp2.scal[1])
        tol = 1e-6 * float(ref.abs().max())
        assert float((out_only - ref).abs().max()) <= tol
        assert torch.equal(out_both, out_only)
        assert float((rec2 - out_only).abs().max()) <= 2.0 ** -21 * 32768.0 * float(sp2.scal[1])
    finally:
        cb.set_precision("fp32")


def test_cached_split_is_not_carried_into_a_graph_capture(cuda):
    """A split (or a published maximum) recorded on a STATIC input by the eager warm-up must not be reused while a
    CUDA graph is captured: the split kernel would be missing from the graph and every replay would convolve the
    warm-up's contents.  Replay on new contents has to equal the eager result on those contents, bit for bit."""
    from creste_public_b200 import ops
    g = torch.Generator().manual_seed(5)
    C, K = 64, 64
    w = torch.randn(K, C, 3, 3, generator=g).to(cuda)
    packed = ops.pack_conv_weight_f16_strided(w)

    def f(x):
        return ops.conv2d_presplit(ops.split_f16_cached(x), packed, K, 3, 3, 1, (1, 1, 1, 1), precision="3xfp16")

    x = torch.randn(2, 16, 24, C, generator=g).to(cuda)          # the static input buffer
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        f(x)                                                      # eager warm-up: records a split on x
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = f(x)
    x2 = torch.randn(2, 16, 24, C, generator=g).to(cuda) * 3.0
    x.copy_(x2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, f(x2.clone()))
