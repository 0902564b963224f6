"""oracle/frame_filter_oracle.py (numpy restatement of the reference's frame -> cloud selections) against hand-built
expectations: the predicates, their thresholds' precision, the crop window orientation, the survivor order and the mask
binarisation are each exercised on a case whose answer can be read off."""
from pathlib import Path

import numpy as np
import pytest

from oracle import frame_filter_oracle as FO


def test_maniskill2_predicates_and_order():
    h = w = 4
    xyzw = np.zeros((h * w, 4), np.float32)
    xyzw[:, 3] = 1.0
    xyzw[:, 2] = 0.1
    xyzw[:, 0] = np.arange(h * w)
    xyzw[3, 3] = 0.0                       # w == 0: dropped
    xyzw[5, 2] = 0.005                     # z == float32(0.005) is NOT > 0.005 in float32: dropped
    xyzw[6, 2] = np.nextafter(np.float32(0.005), np.float32(1))  # just above: kept
    xyzw[7, 2] = -1.0                      # below the ground plane: dropped
    rgb = (np.arange(h * w * 3) % 256).astype(np.uint8).reshape(-1, 3)
    c, col = FO.maniskill2_frame(xyzw, rgb, cam_hw=(h, w))
    keep = [i for i in range(16) if i not in (3, 5, 7)]
    assert c.dtype == np.float32 and col.dtype == np.float32
    assert np.array_equal(c[:, 0], np.array(keep, np.float32))          # ascending original order
    assert np.array_equal(col, rgb[keep].astype(np.float32))
    # include_ground: x > -0.8 replaces z > 0.005
    xyzw[:, 0] -= 1.5                      # x = -1.5, -0.5, 0.5, ...
    c2, _ = FO.maniskill2_frame(xyzw, rgb, cam_hw=(h, w), include_ground=True)
    assert np.array_equal(c2[:, 0], np.array([i - 1.5 for i in range(1, 16) if i != 3], np.float32))


def test_maniskill2_crop_window_is_rows_then_columns():
    h = w = 8
    xyzw = np.ones((2 * h * w, 4), np.float32)  # two cameras
    xyzw[:, 0] = np.arange(2 * h * w)
    rgb = np.zeros((2 * h * w, 3), np.uint8)
    c, _ = FO.maniskill2_frame(xyzw, rgb, cam_hw=(h, w), crop=(2, 1), crop_size=3)
    want = [cam * 64 + r * 8 + col for cam in range(2) for r in range(2, 5) for col in range(1, 4)]
    assert np.array_equal(c[:, 0], np.array(want, np.float32))


def test_rlbench_bounds_are_strict_and_masks_are_binarised():
    b = FO.SCENE_BOUNDS
    pts = np.array([[0.0, 0.0, 1.0],            # inside
                    [b[0], 0.0, 1.0],           # on the lower x bound: strict comparison drops it
                    [np.float32(b[3]), 0.0, 1.0],  # float32(0.7) < 0.7 in float64: kept
                    [0.0, 0.6, 1.0],            # y too large
                    [0.1, 0.1, 0.7],            # inside
                    [0.1, 0.1, 0.5]], np.float32).reshape(1, 2, 3, 3)
    rgb = np.arange(18, dtype=np.float32).reshape(1, 2, 3, 3)
    seg = np.array([5, 7, 201, 9, -3, 2], np.float32).reshape(1, 2, 3)
    c, col = FO.rlbench_frame(pts, rgb, seg)
    assert np.float64(np.float32(b[3])) < b[3]
    assert np.array_equal(c, pts.reshape(-1, 3)[[0, 2, 4]])
    assert np.array_equal(col[:, :3], rgb.reshape(-1, 3)[[0, 2, 4]])
    assert np.array_equal(col[:, 3], np.array([1.0, 0.0, -3.0], np.float32))  # valid id -> 1, invalid id -> 0, negative id kept
    c2, col2 = FO.rlbench_frame(pts, rgb)
    assert col2.shape[1] == 3 and np.array_equal(c2, c)


# ---- pinned against the REFERENCE's own dataset __getitem__ (oracle/gen_golden_frames.py -> tests/golden/frame_filter_ref.npz) ----
GOLD = np.load(Path(__file__).parent / "golden" / "frame_filter_ref.npz")
MS_CASES = ("ms_1cam", "ms_2of3_crop", "ms_ground_crop")
RL_CASES = ("rl_front", "rl_4cam_mask")


@pytest.mark.parametrize("case", MS_CASES)
def test_maniskill2_oracle_matches_reference_getitem(case):
    crop = tuple(int(v) for v in GOLD[f"{case}/crop"])
    c, col = FO.maniskill2_frame(GOLD[f"{case}/xyzw"], GOLD[f"{case}/rgb"], include_ground=bool(GOLD[f"{case}/include_ground"]),
                                 crop=None if crop[0] < 0 else crop)
    ref_c, ref_col = GOLD[f"{case}/out_coord"], GOLD[f"{case}/out_color"]
    assert ref_c.dtype == np.float32 and len(ref_c) > 5000
    assert np.array_equal(c, ref_c)
    assert np.array_equal(col, ref_col.astype(np.float32))


@pytest.mark.parametrize("case", RL_CASES)
def test_rlbench_oracle_matches_reference_getitem(case):
    masks = GOLD[f"{case}/masks"] if bool(GOLD[f"{case}/use_mask"]) else None
    c, col = FO.rlbench_frame(GOLD[f"{case}/point_maps"], GOLD[f"{case}/rgbs"], masks)
    ref_c, ref_col = GOLD[f"{case}/out_coord"], GOLD[f"{case}/out_color"]
    assert ref_c.dtype == np.float64 and len(ref_c) > 500       # the reference keeps float64 here; ToTensorPCD casts later
    assert np.array_equal(c, ref_c.astype(np.float32)) and np.array_equal(c.astype(np.float64), ref_c)
    assert np.array_equal(col, ref_col.astype(np.float32))
    assert col.shape[1] == (4 if masks is not None else 3)
