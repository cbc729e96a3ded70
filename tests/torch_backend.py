"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the C-ABI kernels used by
creste_public_b200.autograd / creste.utils.loss_utils, so that the *structure* of the
differentiable graph (closure of the Functions under first- and second-order differentiation,
BatchNorm composition, loss wiring) can be validated against the reference's own autograd on the
CPU-only build box.  Never imported by the product; the GPU tests run the real kernels."""
import contextlib

import numpy as np
import torch
import torch.nn.functional as F


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def chan_affine(x, a=None, b=None, relu=False):
    y = x if a is None else x * a
    if b is not None:
        y = y + b
    return torch.relu(y) if relu else y.clone()


def relu_bwd(g, y):
    return g * (y > 0).to(g.dtype)


def chan_dot(x, y=None):
    Cc = x.shape[-1]
    v = x if y is None else x * y
    return v.reshape(-1, Cc).sum(0)


def chan_stats(x):
    Cc = x.shape[-1]
    v = x.reshape(-1, Cc).double()
    return torch.stack([v.sum(0), (v * v).sum(0)])


def maxpool2(x):
    return _nhwc(F.max_pool2d(_nchw(x), 2, 2))


def _argmax_mask(x):
    xn = _nchw(x)
    _, idx = F.max_pool2d(xn, 2, 2, return_indices=True)
    return xn, idx


def maxpool2_bwd(x, g):
    xn, idx = _argmax_mask(x)
    dx = torch.zeros_like(xn).flatten(2)
    dx.scatter_(2, idx.flatten(2), _nchw(g).flatten(2))
    return _nhwc(dx.view_as(xn))


def maxpool2_gather(x, gg):
    xn, idx = _argmax_mask(x)
    out = _nchw(gg).flatten(2).gather(2, idx.flatten(2)).view_as(idx)
    return _nhwc(out.to(gg.dtype))


def upsample2(x):
    return _nhwc(F.interpolate(_nchw(x), scale_factor=2, mode="bilinear", align_corners=False))


def upsample2_adjoint(g):
    N, Ho, Wo, Cc = g.shape
    x = torch.zeros(N, Cc, Ho // 2, Wo // 2, dtype=g.dtype, requires_grad=True)
    with torch.enable_grad():
        y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        (dx,) = torch.autograd.grad(y, x, _nchw(g))
    return _nhwc(dx)


def conv_raw(x, w, ph, pw):
    return _nhwc(F.conv2d(_nchw(x), w, padding=(ph, pw)))


def wgrad_raw(x, g, R, S, ph, pw):
    K, Cc = g.shape[-1], x.shape[-1]
    w = torch.zeros(K, Cc, R, S, dtype=x.dtype, requires_grad=True)
    with torch.enable_grad():
        y = F.conv2d(_nchw(x), w, padding=(ph, pw))
        (dw,) = torch.autograd.grad(y, w, _nchw(g))
    return dw


def nchw_to_nhwc(x):
    return _nhwc(x)


def nhwc_to_nchw(x):
    return _nchw(x)


def row_dot(x, y=None, mask=None):
    v = x if y is None else x * y
    if mask is not None:
        v = v * mask.to(v.dtype)
    return v.reshape(x.shape[0], -1).sum(1)


def row_scale(x, s, mask=None):
    v = x * s.view(-1, *([1] * (x.ndim - 1)))
    if mask is not None:
        v = v * mask.to(v.dtype)
    return v


def row_normalize(x, mask=None, eps=1e-5):
    v = x if mask is None else x * mask.to(x.dtype)
    return v / (v.reshape(x.shape[0], -1).sum(1).view(-1, *([1] * (x.ndim - 1))) + eps)


def grad_penalty(G):
    n = G.flatten(2).norm(2, dim=1)
    return ((n - 1) ** 2).mean()


def grad_penalty_bwd(G, g):
    Gd = G.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        p = grad_penalty(Gd)
        (d,) = torch.autograd.grad(p, Gd, g.reshape(()))
    return d


def expert_visitation(traj_rc, map_ds, max_steps, H, W):
    xy = traj_rc / map_ds
    B = xy.shape[0]
    t = torch.linspace(0, 1, max_steps).view(1, 1, -1, 1)
    s, e = xy[:, :-1], xy[:, 1:]
    pts = (s.unsqueeze(2) + t * (e - s).unsqueeze(2)).reshape(B, -1, 2)
    pts = torch.cat([pts, xy[:, -1:]], dim=1)
    r = pts[:, :, 0].clamp(0, H - 1).long()
    c = pts[:, :, 1].clamp(0, W - 1).long()
    cnt = torch.zeros(B, H * W)
    cnt.scatter_add_(1, r * W + c, torch.ones_like(r, dtype=torch.float32))
    cnt[cnt > 1] = 1
    return cnt.view(B, H, W)


def adam_step(p, g, m, v, lr, b1, b2, eps, step, grad_scale=1.0):
    gi = g * grad_scale
    m.lerp_(gi, 1 - b1)
    v.mul_(b2).addcmul_(gi, gi, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, (v.sqrt() / np.sqrt(bc2)).add_(eps), value=-lr / bc1)


@contextlib.contextmanager
def patched():
    """Swap the kernel wrappers for the torch stand-ins (CPU graph-structure tests only)."""
    from creste_public_b200 import autograd as ag
    from creste_public_b200 import ops
    names = ["chan_affine", "relu_bwd", "chan_dot", "chan_stats", "maxpool2", "maxpool2_bwd", "maxpool2_gather",
             "upsample2", "upsample2_adjoint", "nchw_to_nhwc", "nhwc_to_nchw", "row_dot",
             "row_scale", "row_normalize", "grad_penalty", "grad_penalty_bwd", "expert_visitation",
             "adam_step"]
    saved = {n: getattr(ops, n) for n in names}
    saved_ag = (ag._conv_raw, ag._wgrad_raw)
    try:
        for n in names:
            setattr(ops, n, globals()[n])
        ag._conv_raw, ag._wgrad_raw = conv_raw, wgrad_raw
        yield
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)
        ag._conv_raw, ag._wgrad_raw = saved_ag
