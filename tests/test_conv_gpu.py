"""GPU parity tests of the conv family (fp32 FFMA path) against torch CPU fp32 convolutions --
the arithmetic the reference itself runs on CPU.  Tolerance: 2e-5 * max|ref| (fp32 summation
order differs; products are exact fp32)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from creste_public_b200 import ops
    return ops


def _close(a, b, tol=2e-5):
    scale = float(b.abs().max()) + 1e-6
    err = float((a - b).abs().max())
    assert err <= tol * scale, f"err {err:.3e} scale {scale:.3e}"


CASES = [
    # N, C, H, W, K, R, stride, pad(t,b,l,r), act, gate, residual
    (1, 4, 32, 48, 32, 3, 2, (0, 1, 0, 1), "swish", False, False),      # stem, static SAME pad
    (2, 96, 16, 24, 24, 1, 1, (0, 0, 0, 0), "none", True, False),       # MBConv project + SE gate
    (1, 24, 16, 24, 24, 1, 1, (0, 0, 0, 0), "none", True, True),        # project + residual
    (1, 432, 8, 15, 432, 3, 1, (1, 1, 1, 1), "relu", False, False),     # up1-like, K > 128
    (1, 96, 20, 20, 64, 7, 2, (3, 3, 3, 3), "relu", False, False),      # BEV conv1
    (1, 64, 12, 12, 128, 3, 2, (1, 1, 1, 1), "relu", False, False),     # layer2.0.conv1
    (1, 64, 12, 12, 128, 1, 2, (0, 0, 0, 0), "none", False, False),     # downsample
    (1, 128, 9, 11, 6, 1, 1, (0, 0, 0, 0), "none", False, False),       # proj K=6
    (1, 48, 9, 11, 1, 1, 1, (0, 0, 0, 0), "relu", False, False),        # postpool K=1
    (2, 40, 10, 14, 64, 5, 1, (2, 2, 2, 2), "relu", False, False),      # reward prepool 5x5
    (1, 288, 6, 10, 96, 1, 1, (0, 0, 0, 0), "relu", False, False),      # fusion
    (1, 256, 7, 9, 128, 3, 1, (1, 1, 1, 1), "sigmoid", False, True),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nchw", [False, True])
def test_conv_simt_matches_torch(cuda, case, nchw):
    N, C, H, W, K, R, stride, pad, act, use_gate, use_res = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(K, C, R, R, generator=g) / (C * R * R) ** 0.5
    scale = torch.rand(K, generator=g) + 0.5
    shift = torch.randn(K, generator=g) * 0.1
    gate = torch.rand(N, C, generator=g) if use_gate else None
    xin = x * gate.view(N, C, 1, 1) if use_gate else x
    ref = F.conv2d(F.pad(xin, (pad[2], pad[3], pad[0], pad[1])), w, stride=stride)
    ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = torch.randn(ref.shape, generator=g) if use_res else None
    if use_res:
        ref = ref + res
    ref = {"none": lambda t: t, "relu": F.relu, "swish": lambda t: t * torch.sigmoid(t),
           "sigmoid": torch.sigmoid}[act](ref)
    ops = _ops()
    out = ops.conv2d(x.to(cuda).permute(0, 2, 3, 1).contiguous(), ops.pack_conv_weight(w.to(cuda)), K,
                     R, R, stride, pad, scale.to(cuda), shift.to(cuda),
                     gate.to(cuda) if use_gate else None,
                     res.to(cuda).permute(0, 2, 3, 1).contiguous() if use_res else None, act,
                     out_nchw=nchw)
    out = out.cpu() if nchw else out.cpu().permute(0, 3, 1, 2)
    _close(out, ref)


@pytest.mark.parametrize("C,R,stride,pad", [(32, 3, 1, (1, 1, 1, 1)), (96, 3, 2, (0, 1, 0, 1)),
                                             (144, 5, 2, (1, 2, 1, 2)), (1152, 5, 1, (2, 2, 2, 2))])
def test_dwconv_se(cuda, C, R, stride, pad):
    g = torch.Generator().manual_seed(C)
    N, H, W = 2, 14, 18
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(C, 1, R, R, generator=g) / R
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    y = F.conv2d(F.pad(x, (pad[2], pad[3], pad[0], pad[1])), w, stride=stride, groups=C)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    y = y * torch.sigmoid(y)
    ops = _ops()
    out, csum = ops.dwconv_bn_swish(x.to(cuda).permute(0, 2, 3, 1).contiguous(),
                                    w.to(cuda).permute(2, 3, 1, 0).reshape(R * R, C).contiguous(),
                                    scale.to(cuda), shift.to(cuda), R, stride, pad)
    _close(out.cpu().permute(0, 3, 1, 2), y)
    _close(csum.cpu().sum(dim=1), y.sum(dim=(2, 3)), 1e-4)
    Csq = max(1, C // 24)
    wr, br = torch.randn(Csq, C, generator=g) / C ** 0.5, torch.randn(Csq, generator=g) * 0.1
    we, be = torch.randn(C, Csq, generator=g) / Csq ** 0.5, torch.randn(C, generator=g) * 0.1
    m = y.mean(dim=(2, 3))
    s = m @ wr.t() + br
    s = s * torch.sigmoid(s)
    gref = torch.sigmoid(s @ we.t() + be)
    gate = ops.se_gate(csum, y.shape[2] * y.shape[3], wr.to(cuda), br.to(cuda), we.to(cuda), be.to(cuda))
    _close(gate.cpu(), gref, 1e-4)


@pytest.mark.parametrize("scale", [2, 4, (2.0, 153 / 76)])
def test_upsample_concat(cuda, scale):
    g = torch.Generator().manual_seed(3)
    Hi, Wi = (8, 76) if isinstance(scale, tuple) else (6, 10)
    x = torch.randn(2, 12, Hi, Wi, generator=g)
    up = F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False)
    skip = torch.randn(2, 8, up.shape[2], up.shape[3], generator=g)
    ref = torch.cat([skip, up], dim=1)
    ops = _ops()
    out = ops.upsample_concat(skip.to(cuda).permute(0, 2, 3, 1).contiguous(),
                              x.to(cuda).permute(0, 2, 3, 1).contiguous(), up.shape[2:], scale)
    _close(out.cpu().permute(0, 3, 1, 2), ref, 1e-6)
    out = ops.upsample_concat(None, x.to(cuda).permute(0, 2, 3, 1).contiguous(), up.shape[2:], scale)
    _close(out.cpu().permute(0, 3, 1, 2), up, 1e-6)


def test_maxpool_concat_and_layout(cuda):
    g = torch.Generator().manual_seed(4)
    a, b, c = (torch.randn(2, n, 16, 12, generator=g) for n in (32, 6, 2))
    ref = F.max_pool2d(torch.cat([a, b, c], 1), 2, 2)[:, :, :4]
    ops = _ops()
    srcs = [ops.nchw_to_nhwc(t.to(cuda)) for t in (a, b, c)]
    assert torch.equal(srcs[0].cpu(), a.permute(0, 2, 3, 1))
    out, nchw = ops.maxpool2_concat(srcs, rows_out=4, want_nchw=True)
    assert torch.equal(nchw.cpu(), ref)
    assert torch.equal(ops.nhwc_to_nchw(out).cpu(), ref)


@pytest.mark.parametrize("K,C,HW", [(32, 128, (64, 48)), (6, 128, (33, 31)), (2, 128, (16, 16)), (12, 64, (8, 40))])
def test_proj_head_equals_generic_conv_and_transposes(cuda, K, C, HW):
    """creste_proj_head (1x1 conv C -> K <= 32 fused with the NCHW copies) is bit-identical to the exact-fp32 generic
    conv followed by the two layout kernels: same FFMA chain over ascending channels."""
    from creste_public_b200 import ops
    torch.manual_seed(K)
    H, W = HW
    x = torch.randn(3, H, W, C, device="cuda")
    w = torch.randn(K, C, 1, 1, device="cuda") / C ** 0.5
    b = torch.randn(K, device="cuda")
    want = ops.conv2d(x, ops.pack_conv_weight(w), K, 1, 1, 1, (0, 0, 0, 0), None, b, None, None, "none", False, "fp32")
    pred, pred_nchw, x_nchw = ops.proj_head(x, w.reshape(K, C).contiguous(), b)
    assert torch.equal(pred, want)
    assert torch.equal(pred_nchw, ops.nhwc_to_nchw(want))
    assert torch.equal(x_nchw, ops.nhwc_to_nchw(x))
    p2, n2, x2 = ops.proj_head(x, w.reshape(K, C).contiguous(), None, want_nchw=False, want_x_nchw=False)
    assert n2 is None and x2 is None
    assert torch.equal(p2, ops.conv2d(x, ops.pack_conv_weight(w), K, 1, 1, 1, (0, 0, 0, 0), precision="fp32"))
