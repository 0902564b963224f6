"""Seeded synthetic point-cloud batches shared by the CPU and GPU tests."""
from __future__ import annotations

import numpy as np


def clouds(b, n, m, seed=0, kind="uniform", ragged=False):
    """Return xyz (sum n_i, 3) f32, offset (b) i32, new_offset (b) i32.

    kind: "uniform" continuous U([-0.5,0.5]^3) (tie-free with probability 1);
          "lattice" integer lattice / 8 (massive distance ties);
          "dup"     uniform with every point duplicated (zero distances, exact ties).
    """
    rng = np.random.default_rng(seed)
    sizes = rng.integers(max(1, int(0.75 * n)), n + 1, size=b) if ragged else np.full(b, n)
    total = int(sizes.sum())
    if kind == "uniform":
        xyz = rng.uniform(-0.5, 0.5, (total, 3))
    elif kind == "lattice":
        xyz = rng.integers(0, 6, (total, 3)) / 8.0
    elif kind == "dup":
        half = rng.uniform(-0.5, 0.5, ((total + 1) // 2, 3))
        xyz = np.concatenate([half, half])[:total]
        xyz = xyz[rng.permutation(total)]
    else:
        raise ValueError(kind)
    offset = np.cumsum(sizes).astype(np.int32)
    new_sizes = np.minimum(np.full(b, m), np.maximum(sizes, 1)) if m is not None else sizes
    new_offset = np.cumsum(new_sizes).astype(np.int32)
    return xyz.astype(np.float32), offset, new_offset


CASES = [
    # (b, n, m, kind, ragged)
    (2, 512, 256, "uniform", False),
    (4, 1024, 512, "uniform", False),
    (3, 1000, 300, "uniform", True),
    (2, 2048, 1024, "uniform", False),
    (2, 4096, 2048, "uniform", False),
    (3, 700, 128, "lattice", True),
    (2, 1024, 512, "lattice", False),
    (2, 1024, 512, "dup", False),
    (5, 37, 16, "uniform", True),
    (2, 9, 9, "uniform", False),
    (1, 1, 1, "uniform", False),
]
