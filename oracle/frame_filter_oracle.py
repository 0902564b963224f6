"""oracle/frame_filter_oracle.py -- TEST INFRASTRUCTURE (never imported by the product).

numpy restatement of the point selections the reference's dataset classes apply to raw camera frames before the point
cloud transforms (`transform_pcd`: GridSamplePCD ..., oracle/grid_sample_oracle.py):
  * ManiSkill2 (src/data/components/maniskill2/maniskill2_single_task_pcd_act.py:196-224): xyzw of the selected cameras
    as (cams, 128, 128, 4); `rand_crop` zeroes everything outside a 112 x 112 pixel window (rows [x0, x0 + 112), columns
    [y0, y0 + 112)); flatten; keep w > 0; then z > 0.005, or x > -0.8 with `include_ground`;
  * RLBench (src/data/components/rlbench/rlbench_single_task_act.py:266-295): camera maps stacked, cast to float64,
    flattened camera-major; keep the points strictly inside SCENE_BOUNDS (rlbench/constants.py:1); `use_mask`: instance
    ids in `invalid_mask_values` -> 0, remaining ids > 0 -> 1, appended to the colours as a fourth channel.
PINNED: oracle/gen_golden_frames.py runs the reference's UNMODIFIED `__getitem__` methods of both dataset classes (modules
loaded by file path; `h5py` and the `src.utils` package front stubbed, instances made without their file-reading
`__init__`) on seeded synthetic frames and stores what they pass to `transform_pcd` in tests/golden/frame_filter_ref.npz
(5 cases: 1 / 2-of-3 cameras, rand_crop, include_ground, RLBench 1 / 4 cameras with masks, values on the thresholds);
tests/test_frame_filter_cpu.py checks these functions against them bit-exactly, plus hand-built expectations.
"""
from __future__ import annotations

import numpy as np

SCENE_BOUNDS = [-0.3, -0.5, 0.6, 0.7, 0.5, 1.6]  # rlbench/constants.py:1
INVALID_MASK_VALUES = (201, 204, 208, 246)       # rlbench_single_task_act.py:36


def maniskill2_frame(xyzw, rgb, cam_hw=(128, 128), include_ground=False, crop=None, crop_size=112):
    """xyzw (P, 4) f32, rgb (P, 3) (uint8 or float) of the selected cameras, P = cams * h * w.  `crop` = (x0, y0) or None.
    Returns (coord (N, 3) f32, color (N, 3) f32)."""
    h, w = cam_hw
    coords = np.array(xyzw, dtype=np.float32).reshape(-1, h, w, 4)
    if crop is not None:  # :200-208
        x0, y0 = crop
        coords[:, :x0] = 0
        coords[:, x0 + crop_size:] = 0
        coords[:, :, :y0] = 0
        coords[:, :, y0 + crop_size:] = 0
    coords = coords.reshape(-1, 4)
    colors = np.asarray(rgb).reshape(-1, 3)
    colors = colors[coords[..., -1] > 0]  # :215-216
    coords = coords[coords[..., -1] > 0][:, :3]
    if not include_ground:  # :217-223
        colors = colors[coords[..., -1] > 0.005]
        coords = coords[coords[..., -1] > 0.005]
    else:
        colors = colors[coords[..., 0] > -0.8]
        coords = coords[coords[..., 0] > -0.8]
    return coords.astype(np.float32), colors.astype(np.float32)


def rlbench_frame(point_maps, rgbs, masks=None, bounds=SCENE_BOUNDS, invalid_mask_values=INVALID_MASK_VALUES):
    """point_maps / rgbs: (cams, h, w, 3); masks: (cams, h, w) instance ids or None.
    Returns (coord (N, 3) f32, color (N, 3 or 4) f32)."""
    coords = np.stack([np.asarray(p).astype(float) for p in point_maps]).reshape(-1, 3)  # :266-277
    colors = np.stack([np.asarray(c).astype(float) for c in rgbs]).reshape(-1, 3)
    scene_mask = ((coords[:, 0] > bounds[0]) & (coords[:, 0] < bounds[3]) & (coords[:, 1] > bounds[1]) & (coords[:, 1] < bounds[4])
                  & (coords[:, 2] > bounds[2]) & (coords[:, 2] < bounds[5]))  # :278-285
    coords, colors = coords[scene_mask], colors[scene_mask]
    if masks is not None:  # :288-296
        m = np.stack([np.asarray(x).astype(float) for x in masks]).reshape(-1)[scene_mask]
        for v in invalid_mask_values:
            m[m == v] = 0
        m[m > 0] = 1
        colors = np.concatenate([colors, m[:, None]], axis=-1)
    return coords.astype(np.float32), colors.astype(np.float32)
