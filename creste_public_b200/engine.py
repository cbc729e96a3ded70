"""Host-side execution helpers shared by the mirrored modules: BatchNorm folding, weight
packing caches and the global precision policy of the conv family.

Nothing here computes on the CPU at run time: folding and packing are a handful of tiny
tensor ops executed on the GPU once per parameter version, cached, and reused every frame.
"""
import os

import torch

from . import ops

_PRECISION = "fp32"


def set_precision(mode):
    """Conv-family arithmetic: 'fp32' (FFMA, exact products), '3xtf32' (tcgen05 kind::tf32,
    fp32-faithful split), '3xfp16' (tcgen05 kind::f16 on power-of-two-scaled fp16 hi/lo operands:
    the same 11-bit significands as tf32 at twice the tensor-pipe rate), 'tf32' / 'fp16' (single pass: the
    arithmetic class of the reference's own default GPU run, cuDNN TF32; measured error is reported, never
    asserted)."""
    global _PRECISION
    assert mode in ops.PRECISION, mode
    _PRECISION = mode
    ops.PUBLISH_AMAX_MODE = mode


def get_precision():
    return _PRECISION


def require_eval(module):
    if module.training:
        raise NotImplementedError(
            f"{type(module).__name__}: training-mode forward (BatchNorm batch statistics, "
            "drop-connect, autograd) is not implemented in creste_public_b200 yet; call "
            ".eval() -- the reference semantics for inference (running statistics) are what "
            "the sm_100a engine implements.  No PyTorch fallback is provided on purpose.")


# Source of the drop-connect uniforms (efficientnet_pytorch.utils.drop_connect draws
# torch.rand([B,1,1,1]) per residual block).  Tests replace it with a CPU-generator-backed sampler so
# that the masks equal the reference's on the same seed; the default draws on the device.
drop_connect_rand = None


def drop_connect_scale(B, rate, device):
    """Per-sample multiplier floor(keep + U[0,1)) / keep of efficientnet_pytorch's drop_connect."""
    keep = 1.0 - rate
    u = drop_connect_rand(B, device) if drop_connect_rand is not None else torch.rand(B, device=device)
    return (torch.floor(keep + u.float()) / keep).contiguous()


# reductions up to this length run on the exact-fp32 CUDA-core kernels (experiment knob: CRESTE_SIMT_MAX_REDUCTION)
_SIMT_MAX_REDUCTION = int(os.environ.get("CRESTE_SIMT_MAX_REDUCTION", "192"))


def pick_mode(x_shape, K, R, S, stride, pad, mode, prefer_tc=False):
    """Requested precision, or the next stricter one the shape is served by:
    3xfp16 -> 3xtf32 -> fp32 (never a looser one)."""
    x_shape = tuple(int(v) for v in x_shape)
    chain = {"3xfp16": ["3xfp16", "3xtf32"], "3xtf32": ["3xtf32"], "tf32": ["tf32"], "fp16": ["fp16"], "fp32": []}[mode]
    # short reductions (1x1 convs with C <= 192: the MBConv expand / project and head convs) are
    # HBM-bound; measured per shape, the exact-fp32 FFMA kernel beats the tensor-core kernel there
    # (no operand pre-pass, no per-CTA TMEM / barrier set-up): 0.79 vs 1.32 ms at C16->K96 @256x480
    # prefer_tc: the caller feeds this conv a pre-split operand from its producer's epilogue (conv stacks), which
    # removes the pre-pass the measurement above charges to the tensor-core path
    if R * S * x_shape[3] <= _SIMT_MAX_REDUCTION and not (prefer_tc and x_shape[3] >= 64):
        return "fp32"
    for m in chain:
        if ops.tc_supported(x_shape, K, R, S, stride, pad, m):
            return m
    return "fp32"


def _ver(*tensors):
    return tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors if t is not None)


def mark_written(tensors):
    """Tell the pack caches that `tensors` were updated through a raw device pointer (the fused Adam
    step on the flat parameter buffer, the BatchNorm running-statistics update inside
    creste_bn_fwd_finalize): a kernel writing through data_ptr() does not advance torch's version
    counter, and PackCache keys on (data_ptr, _version) -- without this an eval forward after a
    training step would reuse the packed weights / folded BatchNorm factors of the old parameters."""
    ts = [t for t in tensors if t is not None]
    if ts:
        torch._C._increment_version(ts)


class _NoTrace:
    """Suspend torch.jit.trace recording: pack / fold computations run as plain eager code whose RESULTS enter the
    trace as constants -- the same graph whether a cache is cold (first trace run) or warm (the tracer's check run),
    and shapes stay Python ints."""

    def __enter__(self):
        self.state = torch._C._get_tracing_state()
        if self.state is not None:
            torch._C._set_tracing_state(None)

    def __exit__(self, *a):
        if self.state is not None:
            torch._C._set_tracing_state(self.state)


class PackCache:
    """Per-module cache of device-side packed tensors, invalidated by parameter version."""

    def __init__(self):
        self._store = {}

    def get(self, key, tensors, builder):
        v = _ver(*tensors)
        hit = self._store.get(key)
        if hit is not None and hit[0] == v:
            return hit[1]
        with torch.no_grad(), _NoTrace():
            val = builder()
        self._store[key] = (v, val)
        return val


def bn_scale_shift(bn, conv_bias=None):
    """Eval-mode BatchNorm as y = x*scale + shift, with an optional preceding conv bias."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    if conv_bias is not None:
        shift = shift + conv_bias * scale
    return scale.float().contiguous(), shift.float().contiguous()


def carried_amax(t):
    """The max|t| bound that travels with an activation tensor (set by the producing op), or None."""
    return getattr(t, "_amax", None)


def carry_amax(out, *sources):
    """Bound of an op whose outputs are convex combinations / copies of its inputs (bilinear up-sampling, channel
    concat, max-pool): the maximum of the inputs' bounds.  No bound if any input has none."""
    bounds = [carried_amax(s) for s in sources if s is not None]
    if bounds and all(b is not None for b in bounds):
        out._amax = bounds[0] if len(bounds) == 1 else torch.maximum(bounds[0], bounds[1])
    return out


class FusedConv:
    """conv (+bias) (+BN eval) packed for creste_conv2d; built lazily from live parameters."""

    def __init__(self, conv, bn=None, prefer_tc=False):
        self.conv, self.bn, self.prefer_tc = conv, bn, prefer_tc
        self.cache = PackCache()

    def packed(self, mode="fp32"):
        """mode: 'fp32' (SIMT layout), '3xtf32' / 'tf32' (tcgen05 layout, pre-rounded)."""
        conv, bn = self.conv, self.bn
        tensors = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var]
                                              if bn is not None else [])

        def build():
            if mode == "fp32":
                w = ops.pack_conv_weight(conv.weight.detach().float())
            elif mode in ("3xfp16", "fp16"):
                w = ops.pack_conv_weight_f16(conv.weight.detach().float())
            else:
                w = ops.pack_conv_weight_tc(conv.weight.detach().float(), split=(mode == "3xtf32"))
            if bn is not None:
                scale, shift = bn_scale_shift(bn, conv.bias)
            else:
                scale = None
                shift = conv.bias.detach().float().contiguous() if conv.bias is not None else None
            return w, scale, shift
        return self.cache.get("w_" + mode, tensors, build)

    def tc_mode(self, x_shape, pad=None, precision=None):
        """The precision mode this conv would run in on an input of `x_shape` (callers that can hand over a
        pre-split operand ask first)."""
        conv = self.conv
        K, _, R, S = (int(v) for v in conv.weight.shape)      # ints also under torch.jit.trace
        if pad is None:
            ph, pw = conv.padding if isinstance(conv.padding, tuple) else (conv.padding,) * 2
            pad = (ph, ph, pw, pw)
        stride = conv.stride[0] if isinstance(conv.stride, tuple) else conv.stride
        return pick_mode(tuple(x_shape), K, R, S, stride, pad, precision or _PRECISION, self.prefer_tc)

    def out_bound(self, mode):
        """(bound_mul, bound_add) of this conv's output: |out| <= max|x| * bound_mul + bound_add with
        bound_mul = max_k(sum|w_k| * |scale_k|), bound_add = max_k|shift_k| (two host floats, cached per parameter
        version; the 0.1 % head-room covers the rounding of the fp32 accumulation)."""
        conv, bn = self.conv, self.bn
        tensors = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var]
                                              if bn is not None else [])

        def build():
            _, scale, shift = self.packed(mode)
            l1 = conv.weight.detach().float().abs().sum(dim=(1, 2, 3))
            if scale is not None:
                l1 = l1 * scale.abs()
            badd = float(shift.abs().max()) if shift is not None else 0.0
            return float(l1.max()) * 1.001, badd * 1.001
        return self.cache.get("bound_" + mode, tensors, build)

    def split_ok(self, x_shape, gate=None, precision=None):
        """True if this conv, fed an [N,H,W,C] activation of `x_shape`, runs in the fp16 tensor-core mode and can
        therefore take a SplitAct written by its producer's epilogue."""
        mode = precision or _PRECISION
        return (mode in ("3xfp16", "fp16") and gate is None and not torch.jit.is_tracing() and x_shape[-1] % 8 == 0
                and self.tc_mode(x_shape, precision=precision) == mode)

    def __call__(self, x_nhwc, act="none", pad=None, gate=None, residual=None, out_nchw=False,
                 precision=None, split_out=None):
        """split_out: None | "only" | "both" -- ask the epilogue to write the output (also) as the 3xFP16 operand of
        the next tensor-core conv ("only": a SplitAct is returned and the fp32 tensor never exists; "both": the fp32
        tensor is returned with the SplitAct attached as `._split`).  Honoured only when this conv itself runs in the
        fp16 tensor-core mode without a residual; otherwise the plain fp32 tensor is returned."""
        if isinstance(x_nhwc, ops.SplitAct):
            return self._call_presplit(x_nhwc, act, pad, residual, out_nchw, precision, split_out)
        pre = getattr(x_nhwc, "_split", None)
        if pre is not None and gate is None and self.split_ok(x_nhwc.shape, None, precision):
            return self._call_presplit(pre, act, pad, residual, out_nchw, precision, split_out)
        conv = self.conv
        K, _, R, S = (int(v) for v in conv.weight.shape)      # ints also under torch.jit.trace
        if pad is None:
            ph, pw = conv.padding if isinstance(conv.padding, tuple) else (conv.padding,) * 2
            pad = (ph, ph, pw, pw)
        stride = conv.stride[0] if isinstance(conv.stride, tuple) else conv.stride
        mode = precision or _PRECISION
        # shapes the tensor-core kernel does not serve (strided, C = 4 stem, K < 8 heads) run on
        # the exact-fp32 CUDA-core kernel -- a stricter precision, never a looser one
        # (not under torch.jit.trace: the traced `creste::conv2d` op is a pure function of its inputs)
        track = (precision or _PRECISION) in ("3xfp16", "fp16") and not torch.jit.is_tracing()
        mode = pick_mode(tuple(x_nhwc.shape), K, R, S, stride, pad, mode, self.prefer_tc)
        w, scale, shift = self.packed(mode)
        # 3xFP16: max|out| is produced by this conv's epilogue and travels with the tensor (`_amax`), so the next
        # tensor-core conv derives its operand scale from it instead of making an extra amax pass over its input
        amax_out = torch.empty(1, device=x_nhwc.device) if track else None
        so = None
        if (split_out is not None and mode in ("3xfp16", "fp16") and residual is None and not out_nchw and K % 8 == 0
                and not torch.jit.is_tracing()):
            so = (split_out,) + self.out_bound(mode)
        out = ops.conv2d(x_nhwc, w, K, R, S, stride, pad, scale, shift, gate, residual, act, out_nchw, mode,
                         amax_in=carried_amax(x_nhwc) if mode in ("3xfp16", "fp16") else None, amax_out=amax_out,
                         **({"split_out": so} if so is not None else {}))
        return _finish_split(out, so, amax_out if track else None)


def _finish_split(out, so, amax_out):
    """Attach the carried bound / the SplitAct to what ops.conv2d(_presplit) returned."""
    if so is not None and so[0] == "both":
        out, sp = out
        out._split = sp
        sp.amax = amax_out
    if amax_out is not None:
        if isinstance(out, ops.SplitAct):
            out.amax = amax_out
        else:
            out._amax = amax_out
    return out


def _fused_presplit(self, xs, act, pad, residual, out_nchw, precision, split_out=None):
    conv = self.conv
    K, _, R, S = (int(v) for v in conv.weight.shape)
    if pad is None:
        ph, pw = conv.padding if isinstance(conv.padding, tuple) else (conv.padding,) * 2
        pad = (ph, ph, pw, pw)
    stride = conv.stride[0] if isinstance(conv.stride, tuple) else conv.stride
    mode = pick_mode(tuple(xs.shape), K, R, S, stride, pad, precision or _PRECISION, self.prefer_tc)
    if mode not in ("3xfp16", "fp16"):
        raise RuntimeError(f"pre-split operand handed to a conv that runs in mode {mode}")
    w, scale, shift = self.packed(mode)
    amax_out = torch.empty(1, device=xs.device)
    so = None
    if split_out is not None and residual is None and not out_nchw and K % 8 == 0:
        so = (split_out,) + self.out_bound(mode)
    out = ops.conv2d_presplit(xs, w, K, R, S, stride, pad, scale, shift, residual, act, out_nchw, mode, amax_out,
                              split_out=so)
    return _finish_split(out, so, amax_out)


FusedConv._call_presplit = _fused_presplit


def upsample_concat_for(conv, skip, x, out_hw, scale_factor):
    """bilinear(x) (concatenated behind `skip`) as the input of FusedConv `conv`: written directly as that conv's
    3xFP16 operand when the conv runs on the tensor cores and both inputs carry their amax bound; the plain fp32
    tensor (with the bound attached) otherwise."""
    Ct = x.shape[-1] + (0 if skip is None else skip.shape[-1])
    shape = (x.shape[0], out_hw[0], out_hw[1], Ct)
    ax, ask = carried_amax(x), (None if skip is None else carried_amax(skip))
    if (_PRECISION in ("3xfp16", "fp16") and not torch.jit.is_tracing() and Ct % 8 == 0 and ax is not None
            and (skip is None or ask is not None) and conv.tc_mode(shape) == _PRECISION):
        return ops.upsample_concat_split(skip, x, out_hw, scale_factor, ask, ax, want_lo=(_PRECISION == "3xfp16"))
    return carry_amax(ops.upsample_concat(skip, x, out_hw, scale_factor), skip, x)


class GraphedForward:
    """CUDA-graph replay of an eval-mode forward with static input shapes.

    The per-frame path is ~230 kernel launches (+ the ctypes / allocator work around each); at
    B = 1 that host work, not the GPU, bounds the latency.  Capturing the whole forward once and
    replaying it removes it: every launch of libcreste_b200 goes to torch's current stream, all
    scratch comes from torch's capture-aware allocator, packed weights / tensor maps are captured
    by value, and the forward contains no host synchronisation (solve_mdp=False).

        g = GraphedForward(lambda rgbd, p2p: model((rgbd, p2p)), (rgbd, p2p))
        out = g(rgbd_next, p2p_next)        # dict of tensors, overwritten by the next call
    """

    def __init__(self, fn, example_inputs, warmup=3):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                       # fills the pack caches, sets kernel attributes
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """CUDA-graph replay of the forward + losses + backward of a training step with static input shapes.

    A stage-1 step is ~2 500 kernel launches (ours + ATen's) issued from Python through ctypes; at B = 16 that host
    work (~100 ms) exceeds the GPU time of the step (~75 ms).  Capturing zero_grad + forward + losses + backward once
    and replaying it removes it.  What stays eager, by design: the DDP-style buffer broadcast before the step, the
    gradient exchange (ONE flat NCCL all-reduce) and the fused Adam launch after it -- three calls, so the captured
    graph contains no collective and no step-dependent scalar (Adam's bias correction).

        step = GraphedTrainStep(lightning_like_module, example_batch)     # module: _losses(batch), optimizers()
        out = step(batch)                                                 # {"loss": tensor}

    The module's `_losses(batch)` must be free of host synchronisation and data-dependent shapes (stage 1 is; the
    stage-2 contrastive loss samples a data-dependent number of cells and is not).  Warm-up iterations run the same
    forward + backward eagerly on a side stream WITHOUT an optimizer step; BatchNorm running statistics and the RNG
    state are restored afterwards so that the first replayed step starts from the state the caller had."""

    def __init__(self, module, example_batch, warmup=2):
        self.m = module
        self.opt = module.optimizers()
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        bufs = [b for b in module.buffers()]
        saved = [b.clone() for b in bufs]
        rng = torch.cuda.get_rng_state()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        from . import _lib
        with torch.cuda.stream(side):
            for _ in range(warmup):
                n0 = _lib.lib().creste_launch_count()
                self._fwd_bwd()
                self.launches_per_replay = int(_lib.lib().creste_launch_count() - n0)   # our kernels inside the graph
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
        with torch.no_grad():
            for b, s_ in zip(bufs, saved):
                b.copy_(s_)
        torch.cuda.set_rng_state(rng)
        self.opt.zero_grad()

    def _fwd_bwd(self):
        self.opt.zero_grad()
        _, loss_dict, meta, loss = self.m._losses(self.static)
        loss.backward()
        self._logged = ({k: w * v.detach() for k, (w, v) in loss_dict.items()}, {k: v.detach() for k, v in meta.items()})
        return loss.detach()

    def __call__(self, batch):
        from .creste.train_traversability import broadcast_buffers
        with torch.no_grad():
            for k, v in batch.items():
                if torch.is_tensor(v):
                    self.static[k].copy_(v, non_blocking=True)
        broadcast_buffers(self.m.model, self.opt.group)
        self.graph.replay()
        self.opt.step()
        # BatchNorm running statistics were written through raw pointers by the replayed kernels
        mark_written([b for b in self.m.buffers()])
        ld, meta = self._logged
        self.m.logged.update({f"train/{k}": v for k, v in ld.items()})
        self.m.logged.update({f"train/{k}": v for k, v in meta.items()})
        self.m.logged["train/loss"] = self.loss
        return {"loss": self.loss}
