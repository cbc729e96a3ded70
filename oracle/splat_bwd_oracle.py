"""Backward of the bilinear BEV splat -- oracle groundwork for the stage-2 ("next")
training surface.  TEST INFRASTRUCTURE; nothing in the product uses it yet.

Restates, in numpy float64, the gradient of reference creste/models/blocks/splat_projection.py:262-354
(`Camera2MapMulti.splat_soft`, scatter_mode 'mean', min_weight 1.0) w.r.t. the point features and the
(fractional) voxel coordinates -- the formulas a `creste_splat_soft_bwd` kernel has to implement:

  forward   taps (dx,dy) in {0,1}^2:  wX = (1-dx) + (2dx-1) rX,  wY likewise,  w = wX wY,  idx = (Y+dy) W + X+dx
            dens[idx] += w ;  F[c,idx] += w f[c,p] ;  out[c,v] = F[c,v] / max(dens[v], min_weight)
  backward  dF[c,v]  = G[c,v] / max(dens[v], mw)
            dD[v]    = Gd[v] - [dens[v] >= mw] * sum_c G[c,v] F[c,v] / dens[v]^2          (clamp passes at >=)
            df[c,p]  = sum_taps w dF[c,idx]                                                (a gather)
            dw       = sum_c f[c,p] dF[c,idx] + dD[idx]        per valid tap
            dx[p]    = sum_taps (2dx-1) wY dw ,  dy[p] = sum_taps (2dy-1) wX dw            (floor has no gradient)

Pinned against the reference's own autograd on the seeded case of tests/golden/splat_bwd.npz
(tests/test_oracle_cpu.py::test_splat_backward_restatement_matches_reference_golden)."""
import numpy as np


def splat_forward(xy, feats, H, W, min_weight=1.0):
    """xy [N,P,2] float, feats [N,C,P] -> (out [N,C,H*W], dens [N,H*W], F [N,C,H*W]) in float64."""
    xy, feats = np.asarray(xy, np.float64), np.asarray(feats, np.float64)
    N, P, _ = xy.shape
    Cc = feats.shape[1]
    F = np.zeros((N, Cc, H * W))
    dens = np.zeros((N, H * W))
    X, Y = np.floor(xy[..., 0]).astype(np.int64), np.floor(xy[..., 1]).astype(np.int64)
    rX, rY = xy[..., 0] - X, xy[..., 1] - Y
    for dx in (0, 1):
        wX = (1 - dx) + (2 * dx - 1) * rX
        for dy in (0, 1):
            wY = (1 - dy) + (2 * dy - 1) * rY
            X_, Y_ = X + dx, Y + dy
            valid = (X_ >= 0) & (X_ < W) & (Y_ >= 0) & (Y_ < H)
            idx = np.where(valid, Y_ * W + X_, 0)
            w = wX * wY * valid
            for n in range(N):
                np.add.at(dens[n], idx[n], w[n])
                for c in range(Cc):
                    np.add.at(F[n, c], idx[n], w[n] * feats[n, c])
    out = F / np.maximum(dens, min_weight)[:, None, :]
    return out, dens, F


def splat_backward(xy, feats, G, Gd, H, W, min_weight=1.0):
    """Gradients of  sum(out * G) + sum(dens * Gd)  w.r.t. feats [N,C,P] and xy [N,P,2]."""
    xy, feats = np.asarray(xy, np.float64), np.asarray(feats, np.float64)
    G, Gd = np.asarray(G, np.float64), np.asarray(Gd, np.float64)
    out, dens, F = splat_forward(xy, feats, H, W, min_weight)
    dF = G / np.maximum(dens, min_weight)[:, None, :]
    safe = np.where(dens > 0, dens, 1.0)
    dD = Gd - np.where(dens >= min_weight, (G * F).sum(1) / safe ** 2, 0.0)
    X, Y = np.floor(xy[..., 0]).astype(np.int64), np.floor(xy[..., 1]).astype(np.int64)
    rX, rY = xy[..., 0] - X, xy[..., 1] - Y
    N = xy.shape[0]
    dfe = np.zeros_like(feats)
    dxy = np.zeros_like(xy)
    for dx in (0, 1):
        wX = (1 - dx) + (2 * dx - 1) * rX
        for dy in (0, 1):
            wY = (1 - dy) + (2 * dy - 1) * rY
            X_, Y_ = X + dx, Y + dy
            valid = (X_ >= 0) & (X_ < W) & (Y_ >= 0) & (Y_ < H)
            idx = np.where(valid, Y_ * W + X_, 0)
            for n in range(N):
                gF = dF[n][:, idx[n]] * valid[n]                    # [C,P]
                dfe[n] += (wX[n] * wY[n]) * gF
                dw = ((feats[n] * gF).sum(0) + dD[n][idx[n]]) * valid[n]
                dxy[n, :, 0] += (2 * dx - 1) * wY[n] * dw
                dxy[n, :, 1] += (2 * dy - 1) * wX[n] * dw
    return dfe, dxy
