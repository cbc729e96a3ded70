"""CPU tests (-m "not gpu") of the on-disk formats either side of the hot path (creste/datasets/coda_formats.py):
round trips and agreement with the reference's own readers / writers (np.fromfile, cv2, pickle)."""
import pickle

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import synth


def _fmt():
    from creste_public_b200.creste.datasets import coda_formats
    return coda_formats


def test_os1_bin_round_trip(tmp_path):
    f = _fmt()
    pc = np.concatenate([synth.os1_scan(0), np.zeros((131072, 2), np.float32)], axis=1)      # x y z intensity ring
    path = str(tmp_path / "3d_raw_os1_0_10.bin")
    f.write_os1_bin(path, pc)
    back = f.read_os1_bin(path)
    assert back.shape == (f.POINTS_PER_SCAN, f.FEATURES_PER_POINT) and np.array_equal(back, pc)
    # the reference's reader: np.fromfile(...).reshape(POINTS_PER_SCAN, -1)
    assert np.array_equal(np.fromfile(path, dtype=np.float32).reshape(131072, -1), back)
    with pytest.raises(ValueError):
        np.zeros(7, np.float32).tofile(str(tmp_path / "bad.bin"))
        f.read_os1_bin(str(tmp_path / "bad.bin"))


def test_depth_png_round_trip_matches_reference_quantisation(tmp_path):
    """LiDAR sweep -> raster (C oracle of projection.py:64-134) -> uint16-mm PNG -> float32 channel of the input."""
    import cv2
    f = _fmt()
    H, W = 128, 240
    dm, dmm = co.lidar_raster(synth.os1_scan(3)[::4], synth.lidar2camrect(H, W), H, W)
    ref = np.clip(dm * 1000, 0, 65535).astype(np.uint16)             # build_dense_depth.py:461-463
    assert np.array_equal(ref.astype(np.float32), dmm)
    path = str(tmp_path / "depth" / "10.png")
    f.write_depth_png(path, dmm)
    assert np.array_equal(f.read_depth_png(path), dmm)
    assert np.array_equal(cv2.imread(path, -1).astype(np.float32), dmm)       # codapefree_dataloader.py:864-866
    with pytest.raises(ValueError):
        f.write_depth_png(path, dmm + 0.5)
    rgb = (np.random.default_rng(0).random((H, W, 3)) * 255).astype(np.uint8)
    cv2.imwrite(str(tmp_path / "rgb.png"), rgb[..., ::-1])
    x = f.rgbd_from_files(str(tmp_path / "rgb.png"), path)
    assert x.shape == (4, H, W) and np.array_equal(x[3], dmm)
    np.testing.assert_array_equal(x[:3], rgb.transpose(2, 0, 1).astype(np.float32) / 255.0)


def test_counterfactual_pickle_schema(tmp_path):
    f = _fmt()
    expert = synth.expert_poses(2, 50, 256, 256, seed=1)
    cf = synth.counterfactuals(expert)[0]
    path = str(tmp_path / "7" / "120.pkl")
    f.save_counterfactuals(path, cf["trajectories"], cf["rank"], seq=7, frame=120, sample_idx=0)
    raw = pickle.load(open(path, "rb"))                               # what scripts/traversability/rlhf/app.py writes
    assert set(raw) == {"trajectories", "rank", "seq", "frame", "sample_idx"}
    back = f.load_counterfactuals(path)
    assert back["trajectories"].dtype == np.float64 and np.array_equal(back["trajectories"], cf["trajectories"])
    assert np.array_equal(back["rank"], cf["rank"]) and back["seq"] == 7 and back["frame"] == 120
    pickle.dump({"rank": [0]}, open(str(tmp_path / "bad.pkl"), "wb"))
    with pytest.raises(KeyError):
        f.load_counterfactuals(str(tmp_path / "bad.pkl"))
