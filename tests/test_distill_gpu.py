"""GPU parity tests of the stage-1 (distillation) training path (-m gpu).

Kernel level: every entry point of csrc/backbone_train.cu against its plain-PyTorch fp32 statement
(tests/torch_backend.py, evaluated on the CPU).  Step level: one DistillationModel.training_step
through the C ABI against the oracle port (oracle/distill_oracle.py, pinned bit-for-bit to the
reference) and against the golden fixture minted from the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import distill_oracle as do
import torch_backend as tb
from test_distill_cpu import check_golden, compare, ours_step

pytestmark = pytest.mark.gpu


def _t(g, *shape, scale=1.0):
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))


def _close(a, b, rtol=1e-5, what=""):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= rtol * max(ref, 1e-6), (what, err, ref)


@pytest.mark.parametrize("C", [16, 24, 32, 40, 144, 1152])
@pytest.mark.parametrize("act", ["none", "relu", "swish"])
def test_bn_kernels(cuda, C, act):
    from creste_public_b200 import ops
    g = np.random.default_rng(C)
    x, gy = _t(g, 3, 9, 7, C), _t(g, 3, 9, 7, C)
    a, b = _t(g, C), _t(g, C)
    p, q, r = _t(g, C), _t(g, C), _t(g, C)
    _close(ops.chan_moments(x.to(cuda)), tb.chan_moments(x), 1e-12, "moments")
    _close(ops.chan_affine_act(x.to(cuda), a.to(cuda), b.to(cuda), act), tb.chan_affine_act(x, a, b, act), 2e-6, "affine_act")
    gu, sums = ops.bn_act_bwd(gy.to(cuda), x.to(cuda), a.to(cuda), b.to(cuda), act)
    gu0, sums0 = tb.bn_act_bwd(gy, x, a, b, act)
    _close(gu, gu0, 5e-6, "gu")
    _close(sums, sums0, 1e-5, "sums")
    _close(ops.chan_axpby(gy.to(cuda), x.to(cuda), p.to(cuda), q.to(cuda), r.to(cuda)), tb.chan_axpby(gy, x, p, q, r), 2e-6, "axpby")
    # the [C]-sized BatchNorm algebra (one launch each way), running statistics included
    M = x.numel() // C
    w, bb = _t(g, C).abs() + 0.5, _t(g, C)
    rm0, rv0 = _t(g, C), _t(g, C).abs() + 0.5
    rm1, rv1 = rm0.clone().to(cuda), rv0.clone().to(cuda)
    ab0, mi0 = tb.bn_fwd_finalize(tb.chan_moments(x), w, bb, M, 1e-3, 0.01, rm0, rv0)
    ab1, mi1 = ops.bn_fwd_finalize(ops.chan_moments(x.to(cuda)), w.to(cuda), bb.to(cuda), M, 1e-3, 0.01, rm1, rv1)
    _close(ab1, ab0, 1e-6, "ab"); _close(mi1, mi0, 1e-12, "mean_inv"); _close(rm1, rm0, 1e-6, "rm"); _close(rv1, rv0, 1e-6, "rv")
    _close(ops.bn_bwd_finalize(sums, ab1, mi1, M), tb.bn_bwd_finalize(sums0, ab0, mi0, M), 1e-5, "bwd finalize")


@pytest.mark.parametrize("R,stride,H,W", [(3, 1, 10, 14), (3, 2, 10, 14), (5, 1, 8, 12), (5, 2, 8, 12), (5, 2, 4, 6),
                                          (3, 1, 2, 3), (5, 1, 2, 3)])
def test_dwconv_family(cuda, R, stride, H, W):
    from creste_public_b200 import ops
    from creste_public_b200.creste.models.blocks.effnet import same_pad
    g = np.random.default_rng(R * 10 + stride)
    C = 40
    lo, hi = same_pad(R, stride)
    pad = (lo, hi, lo, hi)
    x, w = _t(g, 2, H, W, C), _t(g, R * R, C)
    y0 = tb.dwconv_fwd(x, w, R, stride, pad)
    y = ops.dwconv_fwd(x.to(cuda), w.to(cuda), R, stride, pad)
    assert tuple(y.shape) == tuple(y0.shape)
    _close(y, y0, 2e-6, "fwd")
    gy = _t(g, *y0.shape)
    _close(ops.dwconv_dgrad(gy.to(cuda), w.to(cuda), tuple(x.shape), R, stride, pad),
           tb.dwconv_dgrad(gy, w, tuple(x.shape), R, stride, pad), 2e-6, "dgrad")
    _close(ops.dwconv_wgrad(x.to(cuda), gy.to(cuda), R, stride, pad), tb.dwconv_wgrad(x, gy, R, stride, pad), 1e-5, "wgrad")


def test_se_and_small_kernels(cuda):
    from creste_public_b200 import ops
    g = np.random.default_rng(11)
    B, H, W, C = 3, 6, 5, 672
    x, y = _t(g, B, H, W, C), _t(g, B, H, W, C)
    gate, bvec = _t(g, B, 1, 1, C), _t(g, B, 1, 1, C)
    _close(ops.sample_dot(x.to(cuda), None, 1.0 / (H * W)), tb.sample_dot(x, None, 1.0 / (H * W)), 1e-6, "pool")
    _close(ops.sample_dot(x.to(cuda), y.to(cuda)), tb.sample_dot(x, y), 1e-6, "dot")
    _close(ops.sample_affine(x.to(cuda), gate.to(cuda)), tb.sample_affine(x, gate), 1e-6, "gate")
    _close(ops.sample_affine(None, None, bvec.to(cuda), shape=x.shape), tb.sample_affine(None, None, bvec, shape=x.shape), 0, "bcast")
    for kind in ("swish", "sigmoid"):
        v, gv = _t(g, B, 1, 1, 10), _t(g, B, 1, 1, 10)
        _close(ops.act(v.to(cuda), kind), tb.act(v, kind), 2e-6, kind)
        _close(ops.act_bwd(gv.to(cuda), v.to(cuda), kind), tb.act_bwd(gv, v, kind), 5e-6, kind + "_bwd")
    s = torch.tensor([0.0, 1.25, 1.25])
    _close(ops.add_scaled(x.to(cuda), y.to(cuda), s.to(cuda)), tb.add_scaled(x, y, s), 1e-6, "add_scaled")
    _close(ops.add_scaled(x.to(cuda), y.to(cuda), None), tb.add_scaled(x, y, None), 1e-6, "add")
    _close(ops.chan_slice(x.to(cuda), 112, 320), tb.chan_slice(x, 112, 320), 0, "slice")


@pytest.mark.parametrize("N,H,W,K,pad", [(2, 32, 48, 32, (0, 1, 0, 1)),       # one ragged tile per output row
                                         (1, 18, 300, 32, (0, 1, 0, 1)),      # 150 output columns: 2 full tiles + a tail
                                         (3, 33, 131, 32, (1, 1, 1, 1)),      # odd sizes, padding on every side
                                         (2, 16, 140, 24, (0, 1, 0, 1))])     # warps straddling two taps (K = 24)
def test_stem_wgrad(cuda, N, H, W, K, pad):
    """Weight gradient of the strided C = 4 stem conv (tiled shared-memory kernel) against torch's autograd."""
    from creste_public_b200 import ops
    g = np.random.default_rng(12)
    P = (H + pad[0] + pad[1] - 3) // 2 + 1
    Q = (W + pad[2] + pad[3] - 3) // 2 + 1
    x, gy = _t(g, N, H, W, 4), _t(g, N, P, Q, K)
    _close(ops.wgrad_strided(x.to(cuda), gy.to(cuda), 3, 3, 2, pad), tb.wgrad_strided(x, gy, 3, 3, 2, pad), 1e-5, "stem wgrad")


@pytest.mark.parametrize("K,C,R", [(496, 496, 3), (40, 240, 1), (64, 40, 5), (8, 48, 1)])
def test_pack_weight_f16_kernel(cuda, K, C, R):
    """The one-launch 3xFP16 weight pack against the torch pack the inference path caches: same bytes for the
    contiguous filter, and -- through strides -- for the transposed view of the flipped filter (dgrad)."""
    from creste_public_b200 import ops
    g = np.random.default_rng(K + C)
    w = (_t(g, K, C, R, R) * torch.from_numpy(np.exp(g.uniform(-6, 6, (K, 1, 1, 1))).astype(np.float32))).to(cuda)
    assert torch.equal(ops.pack_conv_weight_f16_strided(w), ops.pack_conv_weight_f16(w))
    wt = w.flip(2, 3).transpose(0, 1)
    assert torch.equal(ops.pack_conv_weight_f16_strided(wt), ops.pack_conv_weight_f16(wt.contiguous()))


def test_wgrad_rows(cuda):
    from creste_public_b200 import ops
    g = np.random.default_rng(14)
    x, gy = _t(g, 4, 1, 1, 1152), _t(g, 4, 1, 1, 48)
    _close(ops.wgrad_rows(x.to(cuda), gy.to(cuda)), tb.wgrad_rows(x, gy), 2e-6, "wgrad_rows")


def test_loss_gradients(cuda):
    from creste_public_b200 import ops
    g = np.random.default_rng(13)
    N, D, H, W = 2, 128, 6, 10
    logits = _t(g, N, D, H, W, scale=3.0)
    lab = torch.from_numpy(g.uniform(300, 25600, (N, H * W)).astype(np.float32))
    lab[0, :7] = 0.0
    lab[1, 3] = 25600.0
    lab[1, 4] = 300.0
    lab[1, 5] = float("nan")
    scale = torch.tensor(0.37)
    _close(ops.ce_depth_bwd(logits.to(cuda), lab.to(cuda), 300.0, 25600.0, scale.to(cuda)),
           tb.ce_depth_bwd(logits, lab, 300.0, 25600.0, scale), 5e-6, "ce bwd")
    pred, gt = _t(g, 2, 1, 16, 6, 10), _t(g, 2, 1, 16, 6, 10)
    gt[0, 0, 3] = float("inf")
    gt[1, 0, 5, 2, 2] = float("-inf")
    _close(ops.masked_mse_bwd(pred.to(cuda), gt.to(cuda), scale.to(cuda)), tb.masked_mse_bwd(pred, gt, scale), 1e-6, "mse bwd")


@pytest.mark.parametrize("N,H,W,C,K,R", [(2, 16, 24, 64, 64, 1), (2, 16, 24, 128, 256, 3), (2, 16, 30, 496, 496, 3),
                                         (2, 20, 28, 112, 72, 3), (3, 17, 23, 72, 200, 1), (1, 32, 60, 432, 432, 3),
                                         (2, 16, 24, 1152, 192, 1), (1, 24, 40, 64, 64, 5), (2, 32, 32, 40, 64, 5),
                                         (2, 32, 32, 32, 16, 1), (2, 32, 32, 48, 8, 1), (2, 24, 24, 16, 96, 1), (1, 32, 32, 144, 24, 1),
                                         # tap groups (several taps per CTA share the g tile): 9 taps = 4 + 4 + 1,
                                         # 25 = 6 x 4 + 1, 128-channel tiles 2 + 2 + 2 + 2 + 1, 7x7 = 12 x 4 + 1
                                         (2, 24, 24, 32, 32, 3), (2, 32, 40, 64, 32, 3), (1, 32, 32, 128, 136, 3),
                                         (1, 32, 32, 24, 40, 7), (2, 16, 24, 96, 64, 3)])
def test_wgrad_tcgen05(cuda, N, H, W, C, K, R):
    """Weight gradient on the tensor cores (MN-major tcgen05 operands, 3xFP16 split) against torch CPU
    float64: ragged pixel boxes, channel tails (C, K not multiples of 64 / 128), 1x1 / 3x3 / 5x5 taps."""
    import torch.nn.functional as F
    from creste_public_b200 import ops
    g = np.random.default_rng(C + K + R)
    pad = (R // 2,) * 4
    x, gy = _t(g, N, H, W, C), _t(g, N, H, W, K, scale=1e-3)
    assert ops.wgrad_tc_supported(tuple(x.shape), K, R, R, pad)
    dw = ops.conv2d_wgrad_tc(x.to(cuda), gy.to(cuda), R, R, pad)
    w = torch.zeros(K, C, R, R, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w, padding=R // 2)
    (ref,) = torch.autograd.grad(y, w, gy.permute(0, 3, 1, 2).double())
    _close(dw, ref, 1e-5, "wgrad_tc")


@pytest.mark.parametrize("precision", ["fp32", "3xfp16"])
def test_training_step_matches_port_and_golden(cuda, golden, precision):
    """One full stage-1 step (forward in train mode, three losses, backward, Adam) on the GPU against the
    oracle port (== the reference bit for bit) on the same seeded case, drop-connect masks included."""
    from creste_public_b200 import engine
    case = do.make_case()
    port = do.port_step(case)
    old = engine.get_precision()
    engine.set_precision(precision)
    try:
        ours = ours_step(case, device=cuda)
    finally:
        engine.set_precision(old)
    # On the boxes measured, 3xfp16 (the default mode) reproduces the reference's ReLU masks on this case (0 of 244
    # tensors over the strict limit) and the exact-fp32 FFMA path flips one element of up3's last ReLU (117 over,
    # profiles/r1c_distill_parity.md).  Which side a |u| ~ 1e-7 pre-activation lands on also depends on the HOST
    # CPU's oneDNN kernels that produce the fp32 yardstick, so the assertion is the flip-robust one for both modes:
    # relative L2 per tensor, and the strict per-tensor limit for at least half of the tensors.
    compare(ours, port, do.port_grads_fp64(case)[0], strict=False)
    check_golden(ours, golden("distill_step.npz"), rtol=2e-4)


def _fp64_with_our_relu_masks(case, ours_relu_out_nhwc, tau):
    """Float64 gradients of the oracle port with every ReLU near-tie (|pre-activation| <= tau of the layer's maximum)
    decided the way the GPU run decided it: pass 0 reads the oracle's masks and requires every disagreement to be
    such a near-tie, pass 1 nudges exactly those inputs across zero.  -> (grads, number of nudged inputs)"""
    ours_mask = [(y > 0).permute(0, 3, 1, 2) for y in ours_relu_out_nhwc]
    st = {"pass": 0, "bump": {}, "n": 0, "seen": 0}

    def hook(i, u):
        if st["pass"] == 1:
            return u + st["bump"][i] if i in st["bump"] else None
        ud = u.detach()
        st["seen"] = i + 1
        assert ud.shape == ours_mask[i].shape, (i, ud.shape, ours_mask[i].shape)
        differ = (ud > 0) != ours_mask[i]
        if differ.any():
            amax, worst = float(ud.abs().max()), float(ud.abs()[differ].max())
            assert worst <= tau * amax, f"ReLU {i}: {int(differ.sum())} masks differ, largest |u64| {worst:.3e} of {amax:.3e}"
            side = torch.where(ours_mask[i], 1.0, -1.0).double()
            st["bump"][i] = (side * (ud.abs() + 1e-3 * tau * amax) - ud) * differ
            st["n"] += int(differ.sum())
        return None

    g = do.port_grads_fp64(case, relu_hook=hook)[0]
    assert st["seen"] == len(ours_mask), (st["seen"], len(ours_mask))
    if st["n"]:
        st["pass"] = 1
        g = do.port_grads_fp64(case, relu_hook=hook)[0]
    return g, st["n"]


@pytest.mark.parametrize("precision", ["fp32", "3xfp16"])
def test_training_step_gradients_with_matched_relu_masks(cuda, precision):
    """Every parameter gradient of the stage-1 step, per tensor, against the float64 oracle whose ReLU near-ties are
    decided like ours -- the strict form of the comparison above: no median, no relative-L2 fallback.  The noise the
    fallback absorbed is discrete (one flipped element of the last ReLUs moves every upstream gradient by ~1e-3,
    profiles/r1c_distill_parity.md); with the ten masks matched what is left is rounding."""
    from creste_public_b200 import engine, ops
    case = do.make_case()
    relu_out = []
    real = ops.chan_affine_act

    def spy(x, a, b, act, **kw):
        y = real(x, a, b, act, **kw)
        if act == "relu":
            relu_out.append(y.detach().cpu())
        return y

    old = engine.get_precision()
    engine.set_precision(precision)
    ops.chan_affine_act = spy
    try:
        ours = ours_step(case, device=cuda)
    finally:
        ops.chan_affine_act = real
        engine.set_precision(old)
    truth, nudged = _fp64_with_our_relu_masks(case, relu_out, tau=1e-4 if precision == "fp32" else 5e-4)
    rows = []
    for k, t in truth.items():
        err, tmax = float(np.abs(ours["grads"][k] - t).max()), float(np.abs(t).max())
        rows.append((err / max(tmax, 1e-6), k, err, tmax))
    rows.sort(reverse=True)
    print(f"[stage-1 step {precision}] {nudged} ReLU inputs nudged in the oracle; worst tensors: "
          + "; ".join(f"{k} {r:.1e} (max {m:.1e})" for r, k, e, m in [q for q in rows if q[3] > 1e-6][:4]))
    bad = [(r, k, e, m) for r, k, e, m in rows if e > 5e-4 * m + 1e-6]
    assert not bad, bad[:10]


def test_training_loss_decreases(cuda):
    """Five steps on one batch: the loss goes down and every parameter stays finite."""
    from creste_public_b200 import configs
    from creste_public_b200.creste.train_pefree import DistillationModel
    case = do.make_case(seed=6, B=2, image_size=(64, 96))
    m = DistillationModel(configs.distill_cfg(case["image_size"]))
    m.model.load_state_dict(case["state_dict"])
    m = m.to(cuda).train()
    inputs = {k: case[k].to(cuda) for k in ("image", "depth_label", "fimg_label")}
    torch.manual_seed(0)
    losses = [float(m.training_step({k: v.clone() for k, v in inputs.items()})["loss"]) for _ in range(5)]
    assert losses[-1] < losses[0], losses
    assert all(torch.isfinite(p).all() for p in m.model.parameters())


@pytest.mark.parametrize("act", ["none", "relu", "swish"])
def test_producers_publish_exact_amax(cuda, act):
    """chan_affine_act / chan_axpby with want_amax: same output bits as without, the published maximum equals
    max|out| exactly, and split_f16 from the published maximum gives the same halves and scale as its own amax pass."""
    from creste_public_b200 import ops
    g = np.random.default_rng(7)
    C = 144
    x, u = _t(g, 2, 11, 13, C, scale=3.0).to(cuda), _t(g, 2, 11, 13, C).to(cuda)
    a, b, r = _t(g, C).to(cuda), _t(g, C).to(cuda), _t(g, C).to(cuda)
    for plain, pub in ((ops.chan_affine_act(x, a, b, act), ops.chan_affine_act(x, a, b, act, want_amax=True)),
                       (ops.chan_axpby(u, x, a, b, r), ops.chan_axpby(u, x, a, b, r, want_amax=True))):
        assert torch.equal(plain, pub)
        am = ops.published_amax(pub)
        assert am is not None and ops.published_amax(plain) is None
        assert float(am) == float(pub.abs().max())
        s_pub, s_own = ops.split_f16(pub), ops.split_f16(plain)
        assert torch.equal(s_pub.hi, s_own.hi) and torch.equal(s_pub.lo, s_own.lo)
        assert torch.equal(s_pub.scal[:2], s_own.scal[:2])
        pub.mul_(2.0)                                    # written again: the record is stale and must not be used
        assert ops.published_amax(pub) is None


@pytest.mark.parametrize("act", ["relu", "swish"])
def test_bn_backward_second_pass_recomputes_gu(cuda, act):
    """chan_axpby_act(g, ...) == chan_axpby(gu, ...) bit for bit, and bn_act_bwd without the gu store returns the
    same sums."""
    from creste_public_b200 import ops
    g = np.random.default_rng(3)
    C = 96
    x, gy = _t(g, 2, 13, 9, C, scale=2.0).to(cuda), _t(g, 2, 13, 9, C).to(cuda)
    a, b, q, r = (_t(g, C).to(cuda) for _ in range(4))
    gu, sums = ops.bn_act_bwd(gy, x, a, b, act)
    none, sums2 = ops.bn_act_bwd(gy, x, a, b, act, want_gu=False)
    assert none is None and torch.equal(sums, sums2)
    two = ops.chan_axpby(gu, x, a, q, r)
    one = ops.chan_axpby_act(gy, x, a, b, act, a, q, r, want_amax=True)
    assert torch.equal(one, two)
    assert float(ops.published_amax(one)) == float(two.abs().max())


def test_upsample_adjoint_reads_a_channel_slice_in_place(cuda):
    from creste_public_b200 import ops
    g = _t(np.random.default_rng(2), 2, 12, 20, 40).to(cuda)
    for c0, Cc, f in ((8, 32, 2), (0, 40, 2), (16, 24, 4)):
        Hi, Wi = 12 // f, 20 // f
        a = ops.upsample_adjoint_slice(g, c0, Cc, Hi, Wi, 1.0 / f)
        b = ops.upsample_adjoint(ops.chan_slice(g, c0, Cc), Hi, Wi, 1.0 / f)
        assert torch.equal(a, b), (c0, Cc, f)


def test_training_step_identical_with_published_amax(cuda):
    """The stage-1 step with the producers' published maxima (no amax passes in front of the tensor-core convs) is
    bit-identical to the step that measures every operand maximum in its own pass."""
    from creste_public_b200 import engine, ops
    case = do.make_case()
    old = engine.get_precision()
    engine.set_precision("3xfp16")
    from creste_public_b200 import autograd as ag
    try:
        a = ours_step(case, device=cuda)
        ops.USE_PUBLISHED_AMAX = False
        ag.FUSED_BN_BWD = False                 # and the two-kernel BatchNorm backward with gu stored
        b = ours_step(case, device=cuda)
    finally:
        ops.USE_PUBLISHED_AMAX = True
        ag.FUSED_BN_BWD = True
        engine.set_precision(old)
    assert a["loss"] == b["loss"]
    assert np.array_equal(a["logits"], b["logits"])
    for k in a["grads"]:
        assert np.array_equal(a["grads"][k], b["grads"][k]), k
