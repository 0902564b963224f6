"""oracle/gen_golden_frames.py -- TEST INFRASTRUCTURE.  Run in the BUILD CONTAINER only (needs /root/reference):

    python -m oracle.gen_golden_frames      # writes tests/golden/frame_filter_ref.npz

Runs the REFERENCE's own dataset `__getitem__` methods on seeded synthetic camera frames and records what they hand to
`transform_pcd` (the filtered coord / color arrays):
  * ManiSkill2GoalPosSingleTaskACTPCDDataset.__getitem__  (src/data/components/maniskill2/maniskill2_single_task_pcd_act.py:175-276)
  * RLBenchSingleTaskACTPCDDataset.__getitem__            (src/data/components/rlbench/rlbench_single_task_act.py:238-377)
The two modules are loaded UNMODIFIED by file path under their real dotted names.  What is stubbed is only what their
`__getitem__` never touches: `h5py` (absent here; the trajectory is handed over already loaded, `cache_traj=True`) and
the `src.utils` package front (its `__init__` pulls in hydra / lightning; the methods use `U.RankedLogger` at import and
`U.io_utils` in `__init__` only).  The instances are made with `object.__new__` + the attributes `__init__` would set, so no
dataset file is needed.  `transform_pcd` is a recorder.  np.random is seeded and the draws `__getitem__` makes
(start_ts, crop_start_x, crop_start_y) are replayed to store the crop window with the case."""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


def _reference_modules():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    for p in ("src", "src.data", "src.data.components", "src.data.components.maniskill2", "src.data.components.rlbench"):
        _pkg(p)
    utils = _pkg("src.utils")

    class RankedLogger:  # stands in for src/utils/pylogger.py (lightning rank-zero logger); unused by __getitem__
        def __init__(self, *a, **k):
            pass

        def info(self, *a, **k):
            pass

    utils.RankedLogger = RankedLogger
    utils.io_utils = types.SimpleNamespace()
    _load("src.utils.rotation_conversions", REF / "src/utils/rotation_conversions.py")
    _load("src.data.components.transformpcd", REF / "src/data/components/transformpcd.py")
    _load("src.data.components.rlbench.constants", REF / "src/data/components/rlbench/constants.py")
    ms = _load("src.data.components.maniskill2.maniskill2_single_task_pcd_act",
               REF / "src/data/components/maniskill2/maniskill2_single_task_pcd_act.py")
    rl = _load("src.data.components.rlbench.rlbench_single_task_act", REF / "src/data/components/rlbench/rlbench_single_task_act.py")
    return ms, rl


class Recorder:
    """transform_pcd stand-in: keeps the dict the dataset built and returns what CollectPCD would (coord + feat tensors)."""

    def __init__(self):
        self.seen = None

    def __call__(self, d):
        self.seen = {k: np.array(v) for k, v in d.items()}
        return dict(coord=torch.from_numpy(np.ascontiguousarray(d["coord"])), feat=torch.from_numpy(np.ascontiguousarray(d["color"])))


def maniskill_frame(seed, cams_total):
    """xyzw (cams_total * 128 * 128, 4) with a mix of w = 0 / w = 1, z around the 0.005 ground cut (exact 0.005 included)
    and x around -0.8; rgb uint8."""
    rng = np.random.default_rng(seed)
    P = cams_total * 128 * 128
    xyz = rng.uniform(-1.0, 1.0, (P, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(-0.01, 0.05, P).astype(np.float32)
    xyz[rng.random(P) < 0.02, 2] = np.float32(0.005)
    xyz[rng.random(P) < 0.02, 0] = np.float32(-0.8)
    w = (rng.random(P) < 0.8).astype(np.float32)
    return np.concatenate([xyz, w[:, None]], axis=1), rng.integers(0, 256, (P, 3)).astype(np.uint8)


def main():
    ms, rl = _reference_modules()
    flat = {}

    # ---- ManiSkill2 -----------------------------------------------------------------------------------------------
    for case, (seed, cams_total, camera_ids, include_ground, rand_crop) in {
        "ms_1cam": (11, 1, (0,), False, False),
        "ms_2of3_crop": (12, 3, (0, 2), False, True),
        "ms_ground_crop": (13, 2, (0, 1), True, True),
    }.items():
        xyzw, rgb = maniskill_frame(seed, cams_total)
        traj = {"actions": np.zeros((1, 8), np.float32),
                "obs": {"agent": {"qpos": np.zeros((1, 9), np.float32)}, "extra": {"goal_pos": np.zeros((1, 3), np.float32)},
                        "pointcloud": {"xyzw": xyzw[None].copy(), "rgb": rgb[None].copy()}}}
        ds = object.__new__(ms.ManiSkill2GoalPosSingleTaskACTPCDDataset)
        rec = Recorder()
        ds.load_count, ds.loop, ds.cache_traj, ds.trajectories = 1, 1, True, [traj]
        ds.camera_ids, ds.point_num_per_cam, ds.include_ground = list(camera_ids), 16384, include_ground
        ds.rand_crop, ds.pointmap, ds.chunk_size, ds.goal_cond_keys = rand_crop, False, 4, ["goal_pos"]
        ds.norm_stats = dict(action_mean=np.zeros(8, np.float32), action_std=np.ones(8, np.float32),
                             qpos_mean=np.zeros(9, np.float32), qpos_std=np.ones(9, np.float32))
        ds.transform_pcd = rec
        np.random.seed(seed)
        ds[0]
        np.random.seed(seed)  # replay the draws of __getitem__: start_ts, then the crop corner
        np.random.choice(1)
        crop = (np.random.randint(0, 128 - 112), np.random.randint(0, 128 - 112)) if rand_crop else (-1, -1)
        sel = np.asarray(camera_ids)
        flat[f"{case}/xyzw"] = xyzw.reshape(cams_total, -1, 4)[sel].reshape(-1, 4)
        flat[f"{case}/rgb"] = rgb.reshape(cams_total, -1, 3)[sel].reshape(-1, 3)
        flat[f"{case}/include_ground"] = np.array(include_ground)
        flat[f"{case}/crop"] = np.array(crop, np.int32)
        flat[f"{case}/out_coord"] = rec.seen["coord"]
        flat[f"{case}/out_color"] = rec.seen["color"]
        print(case, "points", flat[f"{case}/xyzw"].shape[0], "->", rec.seen["coord"].shape[0], "crop", crop,
              rec.seen["coord"].dtype, rec.seen["color"].dtype)

    # ---- RLBench --------------------------------------------------------------------------------------------------
    for case, (seed, cameras, hw, use_mask) in {
        "rl_front": (21, ("front",), 128, False),
        "rl_4cam_mask": (22, ("front", "left_shoulder", "right_shoulder", "wrist"), 64, True),
    }.items():
        rng = np.random.default_rng(seed)
        b = rl.SCENE_BOUNDS
        obs = {"gripper_pose": np.array([0.2, 0.0, 0.9, 0, 0, 0, 1], np.float64), "gripper_open": 1.0, "ignore_collisions": 0.0}
        for cam in cameras:
            pc = rng.uniform(-0.6, 1.8, (hw, hw, 3)).astype(np.float32)
            edge = rng.random((hw, hw)) < 0.03  # points exactly on a bound: the comparison is strict
            pc[edge, 0] = np.float32(b[0])
            edge = rng.random((hw, hw)) < 0.03
            pc[edge, 2] = np.float32(b[5])
            obs[f"{cam}_point_cloud"] = pc
            obs[f"{cam}_rgb"] = rng.integers(0, 256, (hw, hw, 3)).astype(np.uint8)
            obs[f"{cam}_mask"] = rng.choice(np.array([0, 3, 17, 201, 204, 208, 246, 250]), (hw, hw)).astype(np.int32)
        episode = {"demo": [obs, dict(obs)], "task_goal": np.zeros(4, np.float32)}
        ds = object.__new__(rl.RLBenchSingleTaskACTPCDDataset)
        rec = Recorder()
        ds.episodes, ds.cache_episode, ds.loop, ds.root = [("close_jar", episode)], True, 1, ""
        ds.cameras, ds.chunk_size, ds.collision, ds.rot_type = cameras, 2, True, "6d"
        ds.use_mask, ds.invalid_mask_values = use_mask, [201, 204, 208, 246]
        ds.transform_pcd = rec
        np.random.seed(seed)
        ds[0]
        flat[f"{case}/point_maps"] = np.stack([obs[f"{c}_point_cloud"] for c in cameras])
        flat[f"{case}/rgbs"] = np.stack([obs[f"{c}_rgb"] for c in cameras])
        flat[f"{case}/masks"] = np.stack([obs[f"{c}_mask"] for c in cameras])
        flat[f"{case}/use_mask"] = np.array(use_mask)
        flat[f"{case}/out_coord"] = rec.seen["coord"]
        flat[f"{case}/out_color"] = rec.seen["color"]
        print(case, "points", len(cameras) * hw * hw, "->", rec.seen["coord"].shape[0], rec.seen["coord"].dtype, rec.seen["color"].shape)

    np.savez_compressed(OUT / "frame_filter_ref.npz", **flat)
    print("wrote", OUT / "frame_filter_ref.npz", (OUT / "frame_filter_ref.npz").stat().st_size // 1024, "KB")


if __name__ == "__main__":
    main()
