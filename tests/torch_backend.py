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


def _pad4(ph, pw):
    t, b = (ph, ph) if isinstance(ph, int) else ph
    l, r = (pw, pw) if isinstance(pw, int) else pw
    return int(t), int(b), int(l), int(r)


def _padded(xn, ph, pw):
    t, b, l, r = _pad4(ph, pw)
    return F.pad(xn, (l, r, t, b))


def conv_raw(x, w, ph, pw, stride=1):
    return _nhwc(F.conv2d(_padded(_nchw(x), ph, pw), w, stride=stride))


def wgrad_raw(x, g, R, S, ph, pw):
    K, Cc = g.shape[-1], x.shape[-1]
    w = torch.zeros(K, Cc, R, S, dtype=x.dtype, requires_grad=True)
    with torch.enable_grad():
        y = F.conv2d(_padded(_nchw(x), ph, pw), w)
        (dw,) = torch.autograd.grad(y, w, _nchw(g))
    return dw


# ------------------------------------------------------------ stage-2 stand-ins
def dilate(g, stride, Hz, Wz):
    N, P, Q, Cc = g.shape
    z = torch.zeros(N, Hz, Wz, Cc, dtype=g.dtype)
    z[:, : (P - 1) * stride + 1: stride, : (Q - 1) * stride + 1: stride] = g
    return z


def phase_slice(x, stride, a, b, Ha, Wa):
    N, H, W, Cc = x.shape
    out = torch.zeros(N, Ha, Wa, Cc, dtype=x.dtype)
    src = x[:, a::stride, b::stride]
    h, w = min(Ha, src.shape[1]), min(Wa, src.shape[2])
    out[:, :h, :w] = src[:, :h, :w]
    return out


def upsample_adjoint(g, Hi, Wi, ratio):
    N, Ho, Wo, Cc = g.shape
    x = torch.zeros(N, Cc, Hi, Wi, dtype=g.dtype, requires_grad=True)
    with torch.enable_grad():
        y = F.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=False)
        (dx,) = torch.autograd.grad(y, x, _nchw(g))
    return _nhwc(dx)


def upsample_adjoint_slice(g, c0, Cc, Hi, Wi, ratio):
    return upsample_adjoint(g[..., c0:c0 + Cc].contiguous(), Hi, Wi, ratio)


def _frustum(depth, p2p, rng, vox):
    M, Hs, Ws = depth.shape
    u, v = torch.meshgrid(torch.arange(Ws), torch.arange(Hs), indexing="xy")
    cam = torch.stack([u, v, torch.ones_like(u)], 0).unsqueeze(0).to(depth.dtype) * depth.view(M, 1, Hs, Ws)
    cam = torch.cat([cam, torch.ones(M, 1, Hs, Ws, dtype=depth.dtype)], 1)
    xyz = torch.bmm(p2p.to(depth.dtype), cam.flatten(2))[:, :3]                       # [M,3,P]
    lo = torch.tensor(rng[:3], dtype=depth.dtype).view(1, 3, 1)
    hi = torch.tensor(rng[3:], dtype=depth.dtype).view(1, 3, 1)
    mask = ((xyz < hi) & (xyz >= lo)).all(dim=1)
    xy = torch.stack([(-rng[0] - xyz[:, 1]) / vox[0], (-rng[1] - xyz[:, 0]) / vox[1]], dim=-1)
    return xy, xyz[:, 2], mask.to(torch.uint8)


def frustum_to_bev(depth, p2p, rng, vox):
    xy, z, mask = _frustum(depth.detach(), p2p, rng, vox)
    return xy.contiguous(), z.contiguous(), mask


def frustum_bwd(dxy, dz, p2p, shape, voxel):
    # xy / z are affine in depth: differentiate at an arbitrary point
    d = torch.ones(shape, requires_grad=True)
    with torch.enable_grad():
        xy, z, _ = _frustum(d, p2p, [0.0] * 6, list(voxel) + [1.0])
        tot = 0.0
        if dxy is not None:
            tot = tot + (xy * dxy).sum()
        if dz is not None:
            tot = tot + (z * dz).sum()
        (g,) = torch.autograd.grad(tot, d)
    return g


def _splat(xy, feats, mask, H, W, min_weight):
    """feats [N,P,F] -> (bev [N,H,W,F], dens [N,1,H,W]); masked points deposit density only."""
    N, P, Fc = feats.shape
    f = feats if mask is None else feats * mask.view(N, P, 1).to(feats.dtype)
    XY = xy.detach().floor().long()
    r = xy - XY.to(xy.dtype)
    dens = torch.zeros(N, H * W, dtype=xy.dtype)
    vol = torch.zeros(N, H * W, Fc, dtype=xy.dtype)
    for dx in (0, 1):
        wX = (1 - dx) + (2 * dx - 1) * r[..., 0]
        for dy in (0, 1):
            wY = (1 - dy) + (2 * dy - 1) * r[..., 1]
            X_, Y_ = XY[..., 0] + dx, XY[..., 1] + dy
            valid = (X_ >= 0) & (X_ < W) & (Y_ >= 0) & (Y_ < H)
            idx = torch.where(valid, Y_ * W + X_, torch.zeros_like(X_))
            w = wX * wY * valid.to(xy.dtype)
            dens = dens.scatter_add(1, idx, w)
            vol = vol.scatter_add(1, idx.unsqueeze(-1).expand(-1, -1, Fc), w.unsqueeze(-1) * f)
    out = vol / dens.clamp(min=min_weight).unsqueeze(-1)
    return out.view(N, H, W, Fc), dens.view(N, 1, H, W)


def splat_soft(xy, feats_nhwc, mask, H, W, min_weight=1.0, want_nhwc=True, want_nchw=True, want_idx=False):
    N, P, _ = xy.shape
    bev, dens = _splat(xy, feats_nhwc.reshape(N, P, -1), mask, H, W, min_weight)
    return {"bev_nhwc": bev.contiguous(), "bev_nchw": _nchw(bev) if want_nchw else None, "dens": dens, "idx": None}


def splat_soft_bwd(xy, feats_nhwc, mask, bev_nhwc, dens, g_bev, g_dens, min_weight=1.0):
    N, P, _ = xy.shape
    _, H, W, _ = bev_nhwc.shape
    x = xy.detach().clone().requires_grad_(True)
    f = feats_nhwc.detach().reshape(N, P, -1).clone().requires_grad_(True)
    with torch.enable_grad():
        bev, d = _splat(x, f, mask, H, W, min_weight)
        tot = (bev * g_bev).sum()
        if g_dens is not None:
            tot = tot + (d * g_dens).sum()
        gx, gf = torch.autograd.grad(tot, (x, f))
    return gf, gx


def depth_expectation_bwd(logits_nhwc, g_metric, dmin=300.0, dmax=25600.0, out_div=1000.0):
    l = logits_nhwc.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        m = (torch.softmax(l, dim=-1) * torch.linspace(dmin, dmax, l.shape[-1])).sum(-1) / out_div
        (g,) = torch.autograd.grad(m, l, g_metric)
    return g


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


# ------------------------------------------------------------ stage-1 backbone training stand-ins
def _act(u, kind):
    if kind == "relu":
        return torch.relu(u)
    if kind == "swish":
        return u * torch.sigmoid(u)
    if kind == "sigmoid":
        return torch.sigmoid(u)
    return u


def _act_grad(u, kind):
    if kind == "relu":
        return (u > 0).to(u.dtype)
    s = torch.sigmoid(u)
    if kind == "swish":
        return s * (1 + u * (1 - s))
    if kind == "sigmoid":
        return s * (1 - s)
    return torch.ones_like(u)


def chan_moments(x):
    return chan_stats(x)


def bn_fwd_finalize(stats, weight, bias, M, eps, momentum, running_mean, running_var):
    mean = stats[0] / M
    var = (stats[1] / M - mean * mean).clamp_min(0.0)
    if running_mean is not None:
        running_mean.mul_(1 - momentum).add_(mean.float(), alpha=momentum)
        running_var.mul_(1 - momentum).add_((var * (M / max(M - 1, 1))).float(), alpha=momentum)
    inv = torch.rsqrt(var + eps)
    a = inv * (weight.double() if weight is not None else 1.0)
    b = (bias.double() if bias is not None else 0.0) - mean * a
    return torch.stack([a.float(), b.float()]), torch.stack([mean, inv])


def bn_bwd_finalize(sums, ab, mean_inv, M):
    s1, s2, mean, inv, a = sums[0], sums[1], mean_inv[0], mean_inv[1], ab[0].double()
    dgamma = inv * (s2 - mean * s1)
    q = -a * inv * dgamma / M
    r = -a * s1 / M - q * mean
    return torch.stack([dgamma.float(), s1.float(), q.float(), r.float()])


def chan_affine_act(x, a, b, act, want_amax=False):
    return _act(x * a + b, act)


def bn_act_bwd(g, x, a, b, act, want_gu=True):
    gu = g if act in ("none", None) else g * _act_grad(x * a + b, act)
    Cc = x.shape[-1]
    gd, xd = gu.reshape(-1, Cc).double(), x.reshape(-1, Cc).double()
    keep = want_gu or act in ("none", None)
    return (gu if keep else None), torch.stack([gd.sum(0), (gd * xd).sum(0)])


def chan_axpby(u, x, p, q, r, want_amax=False):
    return u * p + x * q + r


def chan_axpby_act(g, x, a, b, act, p, q, r, want_amax=False):
    return (g * _act_grad(x * a + b, act)) * p + x * q + r


def _dw_w(w_rsc, R):
    return w_rsc.view(R, R, 1, -1).permute(3, 2, 0, 1).contiguous()


def dwconv_fwd(x, w_rsc, R, stride, pad):
    pt, pb, pl, pr = pad
    xn = F.pad(_nchw(x), (pl, pr, pt, pb))
    return _nhwc(F.conv2d(xn, _dw_w(w_rsc, R), stride=stride, groups=x.shape[-1]))


def dwconv_dgrad(g, w_rsc, x_shape, R, stride, pad):
    N, H, W, Cc = x_shape
    x = torch.zeros(N, H, W, Cc, dtype=g.dtype, requires_grad=True)
    with torch.enable_grad():
        y = dwconv_fwd(x, w_rsc, R, stride, pad)
        (dx,) = torch.autograd.grad(y, x, g)
    return dx


def dwconv_wgrad(x, g, R, stride, pad):
    w = torch.zeros(R * R, x.shape[-1], dtype=x.dtype, requires_grad=True)
    with torch.enable_grad():
        y = dwconv_fwd(x, w, R, stride, pad)
        (dw,) = torch.autograd.grad(y, w, g)
    return dw


def sample_dot(x, y=None, scale=1.0):
    B, Cc = x.shape[0], x.shape[-1]
    v = x if y is None else x * y
    return (v.reshape(B, -1, Cc).double().sum(1) * scale).float().view(B, 1, 1, Cc)


def sample_affine(x, a=None, b=None, shape=None):
    shape = tuple(x.shape) if x is not None else tuple(shape)
    B, Cc = shape[0], shape[-1]
    out = torch.zeros(shape) if x is None else x.clone()
    bc = (B,) + (1,) * (len(shape) - 2) + (Cc,)
    if a is not None and x is not None:
        out = out * a.reshape(bc)
    if b is not None:
        out = out + b.reshape(bc)
    return out


def act(x, kind):
    return _act(x, kind)


def act_bwd(g, x, kind):
    return g * _act_grad(x, kind)


def add_scaled(inp, x, s=None):
    if s is None:
        return inp + x
    return inp + x * s.view(-1, *([1] * (x.ndim - 1)))


def chan_slice(x, c0, cn):
    return x[..., c0:c0 + cn].contiguous()


def wgrad_strided(x, g, R, S, stride, pad):
    pt, pb, pl, pr = pad
    K, Cc = g.shape[-1], x.shape[-1]
    w = torch.zeros(K, Cc, R, S, dtype=x.dtype, requires_grad=True)
    with torch.enable_grad():
        y = F.conv2d(F.pad(_nchw(x), (pl, pr, pt, pb)), w, stride=stride)
        (dw,) = torch.autograd.grad(y, w, _nchw(g))
    return dw


def wgrad_rows(x, g):
    Cc, K = x.shape[-1], g.shape[-1]
    return (g.reshape(-1, K).t() @ x.reshape(-1, Cc)).view(K, Cc, 1, 1)


def conv2d(x, w_packed, K, R, S, stride=1, pad=(0, 0, 0, 0), scale=None, shift=None, gate=None,
           residual=None, act="none", out_nchw=False, precision="fp32"):
    """Only the raw strided stem call of StemConvFn reaches this stand-in (w_packed = torch weights)."""
    pt, pb, pl, pr = pad
    return _nhwc(F.conv2d(F.pad(_nchw(x), (pl, pr, pt, pb)), w_packed, stride=stride))


def pack_conv_weight(w):
    return w


def upsample_concat(skip, x, out_hw, scale_factor=None, x_first=False):
    if scale_factor is None:
        up = _nhwc(F.interpolate(_nchw(x), size=tuple(out_hw), mode="bilinear", align_corners=False))
    else:
        up = _nhwc(F.interpolate(_nchw(x), scale_factor=scale_factor, mode="bilinear", align_corners=False))
    if skip is None:
        return up
    return torch.cat([up, skip], dim=-1) if x_first else torch.cat([skip, up], dim=-1)


def _bins(label_mm, D, dmin, dmax):
    bin_size = (dmax - dmin) / D
    idx = (label_mm - dmin) / bin_size
    bad = (idx < 0) | (idx > D) | ~torch.isfinite(idx)
    idx = torch.where(bad, torch.full_like(idx, float(D)), idx)
    return idx.long()


def stage1_depth_losses(logits_nchw, pred_bins, label_mm, depth_min, depth_max, beta):
    N, D = logits_nchw.shape[0], logits_nchw.shape[1]
    flat = logits_nchw.reshape(N, D, -1).permute(0, 2, 1)
    gt = _bins(label_mm.reshape(N, -1).float(), D, depth_min, depth_max)
    valid = gt != D
    ce = F.cross_entropy(flat[valid], gt[valid], reduction="sum")
    correct = (flat[valid].argmax(1) == gt[valid]).sum()
    sl = F.smooth_l1_loss(pred_bins.reshape(N, -1)[valid].float(), label_mm.reshape(N, -1)[valid] / 1000.0,
                          beta=beta, reduction="sum")
    return torch.stack([ce.double(), valid.sum().double(), correct.double(), sl.double()])


def ce_depth_bwd(logits_nchw, label_mm, depth_min, depth_max, scale_dev):
    N, D = logits_nchw.shape[0], logits_nchw.shape[1]
    gt = _bins(label_mm.reshape(N, -1).float(), D, depth_min, depth_max)
    valid = (gt != D)
    p = torch.softmax(logits_nchw.reshape(N, D, -1), dim=1)
    onehot = F.one_hot(gt.clamp(max=D - 1), D).permute(0, 2, 1).to(p.dtype)
    d = (p - onehot) * valid.unsqueeze(1).to(p.dtype) * scale_dev.reshape(())
    return d.view_as(logits_nchw)


def masked_mse(pred, gt):
    valid = ~torch.isinf(gt)
    d = (pred - gt)[valid].double()
    return torch.stack([(d * d).sum(), valid.sum().double()])


def masked_mse_bwd(pred, gt, scale_dev):
    valid = ~torch.isinf(gt)
    return torch.where(valid, (pred - gt) * scale_dev.reshape(()), torch.zeros_like(pred))


def depth_expectation(logits_nhwc, dmin=300.0, dmax=25600.0, out_div=1000.0):
    D = logits_nhwc.shape[-1]
    p = torch.softmax(logits_nhwc, dim=-1)
    vals = torch.linspace(dmin, dmax, D)
    return (p * vals).sum(-1) / out_div, logits_nhwc.argmax(-1)


# ------------------------------------------------------------ stage-2 loss stand-ins
def bin_depths(depth, mode, depth_min, depth_max, num_bins, target=False):
    assert mode == "UD"
    idx = (depth - depth_min) / ((depth_max - depth_min) / num_bins)
    if not target:
        return idx
    bad = (idx < 0) | (idx > num_bins) | ~torch.isfinite(idx)
    return torch.where(bad, torch.full_like(idx, float(num_bins)), idx).long()


def _sl1_valid(gt, mask):
    v = torch.isfinite(gt)
    return v if mask is None else v & mask.bool().view_as(gt)


def smooth_l1(pred, gt, mask, gt_scale, beta):
    v = _sl1_valid(gt, mask)
    l = F.smooth_l1_loss(pred[v], gt[v] * gt_scale, beta=beta, reduction="sum")
    return torch.stack([l.double(), v.sum().double()])


def smooth_l1_bwd(pred, gt, mask, gt_scale, beta, scale_dev):
    v = _sl1_valid(gt, mask)
    p = pred.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        l = F.smooth_l1_loss(p[v], gt[v] * gt_scale, beta=beta, reduction="sum")
        (g,) = torch.autograd.grad(l, p)
    return g * scale_dev.reshape(())


def _ce_terms(logits, labels, mask, weights, ignore_index):
    B, Cc = logits.shape[0], logits.shape[1]
    flat = logits.reshape(B, Cc, -1).permute(0, 2, 1).reshape(-1, Cc)
    lab = labels.reshape(-1)
    m = torch.ones_like(lab, dtype=torch.bool) if mask is None else mask.reshape(-1).bool()
    return flat, lab, m


def ce_weighted(logits, labels, mask, weights, ignore_index=-100):
    flat, lab, m = _ce_terms(logits, labels, mask, weights, ignore_index)
    live = m & (lab != ignore_index)
    w = torch.ones(flat.shape[1]) if weights is None else weights
    nll = F.cross_entropy(flat[live], lab[live], weight=w, reduction="sum")
    nz = m & (lab != 0)
    ok = (flat[nz].argmax(1) == lab[nz]).sum()
    return torch.stack([nll.double(), w[lab[live]].sum().double(), ok.double(), nz.sum().double()])


def ce_weighted_bwd(logits, labels, mask, weights, ignore_index, scale_dev):
    l = logits.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        flat, lab, m = _ce_terms(l, labels, mask, weights, ignore_index)
        live = m & (lab != ignore_index)
        w = torch.ones(flat.shape[1]) if weights is None else weights
        nll = F.cross_entropy(flat[live], lab[live], weight=w, reduction="sum")
        (g,) = torch.autograd.grad(nll, l)
    return g * scale_dev.reshape(())


def l2norm_rows(x, eps=1e-12):
    n = x.norm(dim=1).clamp_min(eps)
    return x / n.unsqueeze(1), n


def l2norm_rows_bwd(y, dy, nrm):
    return (dy - y * (y * dy).sum(1, keepdim=True)) / nrm.unsqueeze(1)


def _supcon_rows(f, a, lf, la, self_off, temperature, cw):
    N, Na = f.shape[0], a.shape[0]
    logits = f @ a.t() / temperature
    notself = torch.ones(N, Na, dtype=torch.bool)
    notself[torch.arange(N), torch.arange(N) + self_off] = False
    pos = (lf.view(-1, 1) == la.view(1, -1)) & notself
    logits = logits.masked_fill(~notself, -1e9)
    logp = torch.log_softmax(logits, dim=1)
    p = pos.float() / pos.sum(1, keepdim=True).clamp(min=1.0)
    li = -(p * logp).sum(1)
    if cw is not None:
        li = li * cw[lf]
    return li, logits, pos


def supcon_fwd(f, a, lf, la, self_off, temperature, class_weights=None):
    li, logits, pos = _supcon_rows(f, a, lf, la, self_off, temperature, class_weights)
    m = logits.max(1).values
    stats = torch.stack([m, torch.exp(logits - m.unsqueeze(1)).sum(1), pos.sum(1).float(),
                         (logits * pos).sum(1)], dim=1)
    return stats, li.sum().double().reshape(1)


def supcon_bwd(f, a, lf, la, self_off, temperature, class_weights, stats, scale_dev):
    ff, aa = f.detach().clone().requires_grad_(True), a.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        li, _, _ = _supcon_rows(ff, aa, lf, la, self_off, temperature, class_weights)
        df, da = torch.autograd.grad(li.sum(), (ff, aa))
    return df * scale_dev.reshape(()), da * scale_dev.reshape(())


# ------------------------------------------------------------ eval / IRL stand-ins over the C oracle
def vi_solve(r, gamma=0.99, thr=1e-3, max_sweeps=4096, want_q=True):
    from oracle import c_oracle
    r3 = r.detach().reshape(r.shape[0], r.shape[-2], r.shape[-1]).float().numpy()
    v, q, pi, K = c_oracle.vi_solve(r3, gamma, thr, max_sweeps)
    B, H, W = r3.shape
    return (torch.from_numpy(v).view(B, 1, H, W), torch.from_numpy(q), torch.from_numpy(pi),
            torch.tensor([K, 0], dtype=torch.int32))


def svf(policy, expert_rc, fov, T, ds=2, sharpen=True, temperature=0.005, zero_terminal=False):
    from oracle import c_oracle
    s, st, g = c_oracle.svf(policy.detach().float().numpy(), expert_rc.detach().float().numpy(),
                            fov.detach().to(torch.uint8).numpy(), T, ds, sharpen, temperature, zero_terminal)
    return torch.from_numpy(s), torch.from_numpy(st), torch.from_numpy(g)


def maxpool2_concat(srcs_nhwc, rows_out=None, want_nchw=False):
    x = torch.cat(list(srcs_nhwc), dim=-1)
    y = _nhwc(F.max_pool2d(_nchw(x), 2, 2))
    if rows_out is not None:
        y = y[:, :rows_out].contiguous()
    return (y, _nchw(y)) if want_nchw else y


STAGE1_NAMES = ["bn_fwd_finalize", "bn_bwd_finalize", "chan_moments", "chan_affine_act", "bn_act_bwd", "chan_axpby", "chan_axpby_act", "dwconv_fwd", "dwconv_dgrad",
                "dwconv_wgrad", "sample_dot", "sample_affine", "act", "act_bwd", "add_scaled", "chan_slice",
                "wgrad_strided", "wgrad_rows", "conv2d", "pack_conv_weight", "upsample_concat", "stage1_depth_losses",
                "ce_depth_bwd", "masked_mse", "masked_mse_bwd", "depth_expectation"]
STAGE2_NAMES = ["dilate", "phase_slice", "upsample_adjoint", "upsample_adjoint_slice", "frustum_to_bev", "frustum_bwd", "splat_soft",
                "splat_soft_bwd", "depth_expectation_bwd", "bin_depths", "smooth_l1", "smooth_l1_bwd", "ce_weighted",
                "ce_weighted_bwd", "l2norm_rows", "l2norm_rows_bwd", "supcon_fwd", "supcon_bwd", "vi_solve", "svf",
                "maxpool2_concat"]


@contextlib.contextmanager
def patched():
    """Swap the kernel wrappers for the torch stand-ins (CPU graph-structure tests only)."""
    from creste_public_b200 import autograd as ag
    from creste_public_b200 import ops
    names = ["chan_affine", "relu_bwd", "chan_dot", "chan_stats", "maxpool2", "maxpool2_bwd", "maxpool2_gather",
             "upsample2", "upsample2_adjoint", "nchw_to_nhwc", "nhwc_to_nchw", "row_dot",
             "row_scale", "row_normalize", "grad_penalty", "grad_penalty_bwd", "expert_visitation",
             "adam_step"] + STAGE1_NAMES + STAGE2_NAMES
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
