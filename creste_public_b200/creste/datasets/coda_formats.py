"""On-disk / wire formats either side of the hot path, so the kernels can be fed from
and drained to the reference's own files:

  OS1 sweep `.bin`        131072 points x 5 float32 (x, y, z, intensity, ring) -- creste/datasets/coda_utils.py:3-4,
                          read with np.fromfile(...).reshape(POINTS_PER_SCAN, -1) (codapefree_dataloader.py:781)
  sparse depth `.png`     uint16 millimetres, 0 = no return -- scripts/preprocessing/build_dense_depth.py:461-463
                          (clip to [0, 65535]), read back with cv2.imread(path, -1).astype(float32)
                          (codapefree_dataloader.py:864-866) as channel 3 of the network input
  counterfactual `.pkl`   {trajectories float64 [N,T,2], rank [N] (0 optimal, > 0 sub-optimal), seq, frame,
                          sample_idx} -- scripts/traversability/rlhf/app.py:213-222, consumed by MaxEntIRLLoss through
                          `counterfactuals_label`
  Lightning `.ckpt`       handled by the modules' own load_weights (key surgery in creste/models/*.py)

Host I/O only (numpy / cv2); the arithmetic between these files is the GPU path
(creste_public_b200.ops.lidar_raster: `.bin` sweep -> the exact uint16-mm raster the PNG stores)."""
import os
import pickle

import numpy as np

POINTS_PER_SCAN = 131072
FEATURES_PER_POINT = 5


def read_os1_bin(path, features=None):
    """-> float32 [npts, features]; `features` defaults to whatever divides the file into POINTS_PER_SCAN rows
    (5 for raw CODa sweeps, 4 for the ego-compensated ones build_dense_depth.py:276 reads)."""
    raw = np.fromfile(path, dtype=np.float32)
    if features is None:
        if raw.size % POINTS_PER_SCAN:
            raise ValueError(f"{path}: {raw.size} floats is not a whole number of {POINTS_PER_SCAN}-point rows")
        features = raw.size // POINTS_PER_SCAN
    if features < 3 or raw.size % features:
        raise ValueError(f"{path}: cannot view {raw.size} floats as [*, {features}]")
    return raw.reshape(-1, features)


def write_os1_bin(path, points):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    if pts.ndim != 2 or pts.shape[1] < 3:
        raise ValueError("points must be [n, >= 3]")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    pts.tofile(path)


def write_depth_png(path, depth_mm):
    """uint16-millimetre depth image (already quantised, e.g. ops.lidar_raster's `depth_mm` output)."""
    import cv2
    d = np.asarray(depth_mm)
    if d.dtype != np.uint16:
        if np.any(d < 0) or np.any(d > 65535) or np.any(d != np.floor(d)):
            raise ValueError("depth_mm must hold integers in [0, 65535]")
        d = d.astype(np.uint16)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if not cv2.imwrite(path, d):
        raise IOError(f"cv2.imwrite failed for {path}")


def read_depth_png(path):
    """-> float32 [H,W] millimetres, exactly as the reference loader produces channel 3 of `image`."""
    import cv2
    d = cv2.imread(path, -1)
    if d is None:
        raise IOError(f"cannot read {path}")
    if d.dtype != np.uint16 or d.ndim != 2:
        raise ValueError(f"{path}: expected a single-channel uint16 image, got {d.dtype} {d.shape}")
    return d.astype(np.float32)


def save_counterfactuals(path, trajectories, rank, seq, frame, sample_idx):
    t = np.asarray(trajectories, dtype=np.float64)
    r = np.asarray(rank)
    if t.ndim != 3 or t.shape[2] != 2 or r.shape != (t.shape[0],):
        raise ValueError("trajectories [N,T,2] and rank [N] expected")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        pickle.dump({"trajectories": t, "rank": r, "seq": seq, "frame": frame, "sample_idx": sample_idx}, f)


def load_counterfactuals(path):
    """-> the dict MaxEntIRLLoss reads (`trajectories` float64 [N,T,2], `rank` [N]); validates the schema."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    for k in ("trajectories", "rank"):
        if k not in d:
            raise KeyError(f"{path}: missing '{k}'")
    d["trajectories"] = np.asarray(d["trajectories"], dtype=np.float64)
    d["rank"] = np.asarray(d["rank"])
    if d["trajectories"].ndim != 3 or d["trajectories"].shape[2] != 2 or d["rank"].shape[0] != d["trajectories"].shape[0]:
        raise ValueError(f"{path}: bad shapes {d['trajectories'].shape} / {d['rank'].shape}")
    return d


def rgbd_from_files(rgb_path, depth_png_path):
    """-> float32 [4,H,W]: RGB / 255 + depth in mm, the reference's `_load_rgbd` for one camera without augmentation
    (codapefree_dataloader.py:843-879)."""
    import cv2
    bgr = cv2.imread(rgb_path, -1)
    if bgr is None:
        raise IOError(f"cannot read {rgb_path}")
    rgb = cv2.cvtColor(bgr.astype(np.uint8), cv2.COLOR_BGR2RGB).astype(np.float32).transpose(2, 0, 1) / 255.0
    depth = read_depth_png(depth_png_path)[None]
    if depth.shape[1:] != rgb.shape[1:]:
        raise ValueError(f"image {rgb.shape[1:]} and depth {depth.shape[1:]} sizes differ")
    return np.concatenate([rgb, depth], axis=0)

