"""oracle/gen_golden_grid.py -- TEST INFRASTRUCTURE.  Run in the BUILD CONTAINER only (needs /root/reference):

    python -m oracle.gen_golden_grid      # writes tests/golden/grid_sample_ref.npz

Runs the REFERENCE's own `GridSamplePCD` (src/data/components/transformpcd.py:664-793, loaded by file path -- the module
needs numpy and torch only), `NormalizeColorPCD`, `ToTensorPCD`, `CollectPCD` and `pcd_collate_fn`
(src/utils/sparse_tensor_utils.py:65-82) in the configuration of configs/data/maniskill2_act_pcd_dataset.yaml:15-34
(grid_size 0.005, hash fnv, return_grid_coord, keys [coord, color], feat_keys [color, coord]) in TEST mode (part 0) on
seeded synthetic raw clouds, and stores inputs + outputs.  The stored `grid_coord` / `offset` do not depend on the
reference's unstable argsort; `coord` / `feat` are stored too and compared only through properties (every output point
is a member of the voxel it represents)."""
from __future__ import annotations

import importlib.util
import sys
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


def raw_clouds(seed, b, n, scale):
    rng = np.random.default_rng(seed)
    clouds = []
    for i in range(b):
        ni = int(rng.integers(int(0.6 * n), n + 1))
        centre = rng.uniform(-0.3, 0.3, 3)
        coord = (centre + scale * rng.standard_normal((ni, 3))).astype(np.float32)  # dense blob: many points per voxel
        color = rng.integers(0, 256, (ni, 3)).astype(np.float32)
        clouds.append((coord, color))
    return clouds


def main():
    T = _load("ref_transformpcd", REF / "src" / "data" / "components" / "transformpcd.py")
    S = _load("ref_sparse_tensor_utils", REF / "src" / "utils" / "sparse_tensor_utils.py")
    flat = {}
    for case, (seed, b, n, scale, gs) in {"dense": (1, 4, 3000, 0.03, 0.005), "sparse": (2, 3, 800, 0.5, 0.005),
                                          "coarse": (3, 2, 2000, 0.2, 0.05)}.items():
        clouds = raw_clouds(seed, b, n, scale)
        samples = []
        for coord, color in clouds:
            gsamp = T.GridSamplePCD(grid_size=gs, hash_type="fnv", mode="test", return_grid_coord=True, keys=("coord", "color"))
            part0 = gsamp(dict(coord=coord.copy(), color=color.copy()))[0]
            d = T.NormalizeColorPCD()(dict(part0))
            d = T.ToTensorPCD()(d)
            d = T.CollectPCD(keys=("coord", "grid_coord"), feat_keys=("color", "coord"))(d)
            samples.append({"pcds": [d], "qpos": torch.zeros(1)})
        batch = S.pcd_collate_fn(samples)["pcds"]
        flat[f"{case}/grid_size"] = np.array(gs)
        flat[f"{case}/sizes"] = np.array([c.shape[0] for c, _ in clouds])
        flat[f"{case}/in_coord"] = np.concatenate([c for c, _ in clouds])
        flat[f"{case}/in_color"] = np.concatenate([c for _, c in clouds]).astype(np.uint8)
        for k in ("coord", "grid_coord", "feat", "offset"):
            flat[f"{case}/out_{k}"] = batch[k].numpy()
        print(case, "raw", flat[f"{case}/in_coord"].shape[0], "->", batch["coord"].shape[0], "offset", batch["offset"].tolist())
    np.savez_compressed(OUT / "grid_sample_ref.npz", **flat)
    print("wrote", OUT / "grid_sample_ref.npz", (OUT / "grid_sample_ref.npz").stat().st_size // 1024, "KB")


if __name__ == "__main__":
    main()
