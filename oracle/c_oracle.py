"""ctypes binding of oracle/creste_oracle.c.  TEST INFRASTRUCTURE (checker only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (creste_public_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcreste_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "creste_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def num_threads():
    return int(lib().oracle_num_threads())


def vi_solve(r, gamma=0.99, thr=1e-3, max_sweeps=100000, want_q=True):
    """r [B,H,W] or [B,1,H,W] float32 -> v [B,H,W], q [B,8,H,W], pi [B,8,H,W], sweeps."""
    r = _f32(r)
    if r.ndim == 4:
        r = r[:, 0]
    r = np.ascontiguousarray(r)
    B, H, W = r.shape
    v = np.empty_like(r)
    q = np.empty((B, 8, H, W), np.float32) if want_q else None
    pi = np.empty((B, 8, H, W), np.float32) if want_q else None
    k = C.c_int(0)
    rc = lib().oracle_vi_solve(_p(r, C.c_float), _p(v, C.c_float), _p(q, C.c_float),
                               _p(pi, C.c_float), B, H, W, C.c_float(gamma), C.c_float(thr),
                               max_sweeps, C.byref(k))
    assert rc == 0
    return v, q, pi, k.value


def svf(policy, expert_rc, fov, T, ds=2, sharpen=True, temperature=0.005, zero_terminal=False):
    """policy [B,8,H,W]; expert_rc [B,T,2] (expert[:,:,:2,2]); fov [H,W] bool."""
    policy = _f32(policy)
    expert_rc = _f32(expert_rc)
    fov = np.ascontiguousarray(np.asarray(fov, dtype=np.uint8))
    B, A, H, W = policy.shape
    assert A == 8 and expert_rc.shape == (B, T, 2) and fov.shape == (H, W)
    out = np.empty((B, H, W), np.float32)
    states = np.empty((B, T, 2), np.int64)
    grid = np.empty((B, H, W), np.float32)
    rc = lib().oracle_svf(_p(policy, C.c_float), _p(expert_rc, C.c_float), _p(fov, C.c_uint8), B, H,
                          W, T, ds, int(bool(sharpen)), C.c_float(temperature), int(zero_terminal),
                          _p(out, C.c_float), _p(states, C.c_int64), _p(grid, C.c_float))
    assert rc == 0
    return out, states, grid


def frustum_to_bev(depth, p2p, pc_range, voxel):
    """depth [N,Hs,Ws]; p2p [N,4,4] -> xyz [N,3,P], xy [N,P,2], mask [N,P] (bool)."""
    depth = _f32(depth)
    p2p = _f32(p2p)
    N, Hs, Ws = depth.shape
    P = Hs * Ws
    xyz = np.empty((N, 3, P), np.float32)
    xy = np.empty((N, P, 2), np.float32)
    mask = np.empty((N, P), np.uint8)
    rng = _f32(pc_range)
    vox = _f32(voxel)
    rc = lib().oracle_frustum_to_bev(_p(depth, C.c_float), _p(p2p, C.c_float), _p(rng, C.c_float),
                                     _p(vox, C.c_float), N, Hs, Ws, _p(xyz, C.c_float),
                                     _p(xy, C.c_float), _p(mask, C.c_uint8))
    assert rc == 0
    return xyz, xy, mask.astype(bool)


def splat_soft(xy, feats, H, W, min_weight=1.0):
    """xy [N,P,2]; feats [N,F,P] -> vol [N,F,H*W], dens [N,H*W], idx [N,P,4], wts [N,P,4]."""
    xy = _f32(xy)
    feats = _f32(feats)
    N, P, _ = xy.shape
    F = feats.shape[1]
    vol = np.empty((N, F, H * W), np.float32)
    dens = np.empty((N, H * W), np.float32)
    idx = np.empty((N, P, 4), np.int64)
    wts = np.empty((N, P, 4), np.float32)
    rc = lib().oracle_splat_soft(_p(xy, C.c_float), _p(feats, C.c_float), N, P, F, H, W,
                                 C.c_float(min_weight), _p(vol, C.c_float), _p(dens, C.c_float),
                                 _p(idx, C.c_int64), _p(wts, C.c_float))
    assert rc == 0
    return vol, dens, idx, wts


def lidar_raster(pc, P34, H, W):
    """pc [n,>=3] float32; P34 [3,4] float64 -> depth_m [H,W] f32, depth_mm [H,W] f32."""
    pc = _f32(pc)
    P34 = np.ascontiguousarray(np.asarray(P34, dtype=np.float64))
    dm = np.empty((H, W), np.float32)
    dmm = np.empty((H, W), np.float32)
    rc = lib().oracle_lidar_raster(_p(pc, C.c_float), pc.shape[0], pc.shape[1], _p(P34, C.c_double),
                                   H, W, _p(dm, C.c_float), _p(dmm, C.c_float))
    assert rc == 0
    return dm, dmm


def depth_expectation(logits, dmin=300.0, dmax=25600.0):
    """logits [N,D,Hs,Ws] -> metric [N,Hs,Ws] (m), bins [N,Hs,Ws] int64."""
    logits = _f32(logits)
    N, D, Hs, Ws = logits.shape
    metric = np.empty((N, Hs, Ws), np.float32)
    bins = np.empty((N, Hs, Ws), np.int64)
    rc = lib().oracle_depth_expectation(_p(logits, C.c_float), N, D, Hs * Ws, C.c_float(dmin),
                                        C.c_float(dmax), _p(metric, C.c_float), _p(bins, C.c_int64))
    assert rc == 0
    return metric, bins


def expert_visitation(rc_traj, map_ds, H, W, is_f64):
    """rc_traj [B,T,2] -> counts [B,H,W] in {0,1}."""
    t = np.ascontiguousarray(np.asarray(rc_traj, dtype=np.float64))
    B, T, _ = t.shape
    counts = np.empty((B, H, W), np.float32)
    rc = lib().oracle_expert_visitation(_p(t, C.c_double), B, T, C.c_double(map_ds), H, W,
                                        int(is_f64), _p(counts, C.c_float))
    assert rc == 0
    return counts
