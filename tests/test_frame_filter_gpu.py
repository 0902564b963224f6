"""csrc/frame_filter.cu (through the C ABI / data_gpu.filter_frames_*) against oracle/frame_filter_oracle.py: BIT-EXACT
survivor sets, order, coordinates, colours and per-sample offsets for both dataset families, incl. the random-crop window,
include_ground, uint8 and fp32 colours, instance masks, samples with no survivors, and values sitting on the thresholds;
then the composition with the voxel-grid stage (frames -> grid_sample_collate) against the oracles' composition."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import frame_filter_oracle as FO
from oracle import grid_sample_oracle as GO

pytestmark = pytest.mark.gpu


def _ms2_frames(b, cams, seed):
    rng = np.random.default_rng(seed)
    P = cams * 128 * 128
    xyzw = rng.uniform(-1.0, 1.0, (b, P, 4)).astype(np.float32)
    xyzw[..., 3] = (rng.uniform(0, 1, (b, P)) > 0.3).astype(np.float32)       # 30 % invalid pixels
    xyzw[..., 2] = rng.uniform(-0.01, 0.3, (b, P)).astype(np.float32)
    thr = rng.integers(0, P, 200)
    xyzw[:, thr, 2] = np.float32(0.005)                                          # exactly on the ground threshold
    xyzw[:, thr[:50], 0] = np.float32(-0.8)
    rgb = rng.integers(0, 256, (b, P, 3)).astype(np.uint8)
    return xyzw, rgb


@pytest.mark.parametrize("cams,include_ground,use_crop,u8", [(1, False, False, True), (2, True, False, False), (2, False, True, True),
                                                             (4, True, True, True)])
def test_maniskill2_frames_bit_exact(cams, include_ground, use_crop, u8):
    from pointcloudmatters_b200.data_gpu import filter_frames_maniskill2

    b = 3
    xyzw, rgb = _ms2_frames(b, cams, seed=cams)
    xyzw[1, :, 3] = 0.0  # a sample with no surviving point
    crop = np.array([[3, 9], [0, 15], [15, 0]], np.int32) if use_crop else None
    col_in = rgb if u8 else rgb.astype(np.float32) * 0.5
    c, col, off = filter_frames_maniskill2(torch.from_numpy(xyzw).cuda(), torch.from_numpy(col_in).cuda(), include_ground=include_ground,
                                           crop=crop)
    want_c, want_col, sizes = [], [], []
    for i in range(b):
        wc, wcol = FO.maniskill2_frame(xyzw[i], col_in[i], include_ground=include_ground, crop=None if crop is None else tuple(crop[i]))
        want_c.append(wc); want_col.append(wcol); sizes.append(len(wc))
    assert sizes[1] == 0 and sizes[0] > 1000
    assert np.array_equal(off.cpu().numpy(), np.cumsum(sizes))
    assert np.array_equal(c.cpu().numpy(), np.concatenate(want_c))
    assert np.array_equal(col.cpu().numpy(), np.concatenate(want_col))


@pytest.mark.parametrize("with_masks", [False, True])
def test_rlbench_frames_bit_exact(with_masks):
    from pointcloudmatters_b200.data_gpu import filter_frames_rlbench

    rng = np.random.default_rng(5)
    b, cams, h, w = 2, 4, 128, 128
    lo, hi = np.array(FO.SCENE_BOUNDS[:3]) - 0.2, np.array(FO.SCENE_BOUNDS[3:]) + 0.2
    pts = rng.uniform(lo, hi, (b, cams, h, w, 3)).astype(np.float32)
    edge = rng.integers(0, h * w, 300)
    flat = pts.reshape(b, cams, -1, 3)
    flat[:, :, edge[:100], 0] = np.float32(FO.SCENE_BOUNDS[0])   # on the bounds, as float32
    flat[:, :, edge[100:200], 1] = np.float32(FO.SCENE_BOUNDS[4])
    flat[:, :, edge[200:], 2] = np.float32(FO.SCENE_BOUNDS[5])
    rgb = rng.integers(0, 256, (b, cams, h, w, 3)).astype(np.float32)
    seg = rng.choice(np.array([0, 3, 17, 201, 204, 208, 246, 250, -1], np.float32), (b, cams, h, w)) if with_masks else None
    c, col, off = filter_frames_rlbench(torch.from_numpy(pts).cuda(), torch.from_numpy(rgb).cuda(),
                                        None if seg is None else torch.from_numpy(seg).cuda())
    want = [FO.rlbench_frame(pts[i], rgb[i], None if seg is None else seg[i]) for i in range(b)]
    assert np.array_equal(off.cpu().numpy(), np.cumsum([len(x[0]) for x in want]))
    assert np.array_equal(c.cpu().numpy(), np.concatenate([x[0] for x in want]))
    assert np.array_equal(col.cpu().numpy(), np.concatenate([x[1] for x in want]))
    assert col.shape[1] == (4 if with_masks else 3)


def test_frames_to_collated_batch_matches_the_oracle_pipeline():
    """frames -> filter -> voxel-grid subsampling + collation, all on the device, against the two oracles chained."""
    from pointcloudmatters_b200.data_gpu import filter_frames_maniskill2, grid_sample_collate

    b = 4
    xyzw, rgb = _ms2_frames(b, 1, seed=11)
    c, col, off = filter_frames_maniskill2(torch.from_numpy(xyzw).cuda(), torch.from_numpy(rgb).cuda())
    got = grid_sample_collate(c, col, off, grid_size=0.02, mode="test")
    clouds = [FO.maniskill2_frame(xyzw[i], rgb[i]) for i in range(b)]
    want = GO.grid_sample_collate(clouds, 0.02)
    for k in ("coord", "grid_coord", "feat", "offset"):
        assert np.array_equal(got[k].cpu().numpy(), want[k]), k


# ---- against what the REFERENCE's own dataset __getitem__ produced (tests/golden/frame_filter_ref.npz, oracle/gen_golden_frames.py) ----
GOLD = np.load(Path(__file__).parent / "golden" / "frame_filter_ref.npz")


@pytest.mark.parametrize("case", ["ms_1cam", "ms_2of3_crop", "ms_ground_crop"])
def test_maniskill2_frames_match_reference_getitem(case):
    from pointcloudmatters_b200.data_gpu import filter_frames_maniskill2

    crop = GOLD[f"{case}/crop"]
    c, col, off = filter_frames_maniskill2(torch.from_numpy(GOLD[f"{case}/xyzw"])[None].cuda(), torch.from_numpy(GOLD[f"{case}/rgb"])[None].cuda(),
                                           include_ground=bool(GOLD[f"{case}/include_ground"]), crop=None if crop[0] < 0 else crop[None])
    assert int(off[-1]) == len(GOLD[f"{case}/out_coord"])
    assert np.array_equal(c.cpu().numpy(), GOLD[f"{case}/out_coord"])
    assert np.array_equal(col.cpu().numpy(), GOLD[f"{case}/out_color"].astype(np.float32))


@pytest.mark.parametrize("case", ["rl_front", "rl_4cam_mask"])
def test_rlbench_frames_match_reference_getitem(case):
    from pointcloudmatters_b200.data_gpu import filter_frames_rlbench

    use_mask = bool(GOLD[f"{case}/use_mask"])
    masks = torch.from_numpy(GOLD[f"{case}/masks"].astype(np.float32))[None].cuda() if use_mask else None
    c, col, off = filter_frames_rlbench(torch.from_numpy(GOLD[f"{case}/point_maps"])[None].cuda(),
                                        torch.from_numpy(GOLD[f"{case}/rgbs"])[None].cuda(), masks)
    assert int(off[-1]) == len(GOLD[f"{case}/out_coord"])
    assert np.array_equal(c.cpu().numpy().astype(np.float64), GOLD[f"{case}/out_coord"])
    assert np.array_equal(col.cpu().numpy(), GOLD[f"{case}/out_color"].astype(np.float32))
