"""Pins oracle/grid_sample_oracle.py (numpy restatement of GridSamplePCD + NormalizeColorPCD + CollectPCD + pcd_collate_fn)
to outputs of the REFERENCE classes themselves (tests/golden/grid_sample_ref.npz, oracle/gen_golden_grid.py)."""
import os

import numpy as np
import pytest

from oracle import grid_sample_oracle as G

GOLD = os.path.join(os.path.dirname(__file__), "golden", "grid_sample_ref.npz")
CASES = ["dense", "sparse", "coarse"]


def load_case(g, case):
    sizes = g[f"{case}/sizes"]
    ends = np.cumsum(sizes)
    coord, color = g[f"{case}/in_coord"], g[f"{case}/in_color"].astype(np.float32)
    clouds = [(coord[e - s:e], color[e - s:e]) for s, e in zip(sizes, ends)]
    return clouds, float(g[f"{case}/grid_size"])


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_grid_sample(case):
    g = np.load(GOLD)
    clouds, gs = load_case(g, case)
    got = G.grid_sample_collate(clouds, gs)
    # independent of the reference's unstable argsort: voxel sequence, voxel counts, offsets -- bit-exact
    assert np.array_equal(got["offset"], g[f"{case}/out_offset"])
    assert np.array_equal(got["grid_coord"], g[f"{case}/out_grid_coord"])
    # the member the reference kept lies in the same voxel as ours (and equals ours wherever the voxel has one point)
    ref_coord, ref_feat = g[f"{case}/out_coord"], g[f"{case}/out_feat"]
    start = 0
    for (coord, color), end in zip(clouds, got["offset"]):
        grid_all = np.floor(coord / np.array(gs)).astype(int)
        gmin = grid_all.min(0)
        ref_grid = np.floor(ref_coord[start:end] / np.array(gs)).astype(int) - gmin
        assert np.array_equal(ref_grid, got["grid_coord"][start:end])
        start = end
    single = np.all(np.isclose(ref_coord, got["coord"]), axis=1)
    assert single.mean() > 0.3
    np.testing.assert_array_equal(ref_feat[single], got["feat"][single])  # color / 127.5 - 1 and [color, coord] layout


def test_fnv_hash_known_values():
    # FNV64-1A over three uint64 words (transformpcd.py:775-793), computed by hand in Python integers
    def fnv(v):
        h = 14695981039346656037
        for x in v:
            h = (h * 1099511628211) % 2 ** 64
            h ^= x
        return h

    arr = np.array([[0, 0, 0], [1, 2, 3], [255, 65535, 2 ** 31 - 1]], dtype=np.int64)
    assert [int(x) for x in G.fnv_hash_vec(arr)] == [fnv(r) for r in arr.tolist()]


def test_priority_rule_and_ordering():
    rng = np.random.default_rng(0)
    coord = rng.uniform(0, 0.05, (500, 3)).astype(np.float32)
    idx, grid = G.grid_sample(coord, 0.01)
    key = G.fnv_hash_vec(grid)
    assert np.all(key[1:] > key[:-1])  # ascending, unique keys
    # default priority: the lowest index of every voxel
    gall = np.floor(coord / 0.01).astype(int)
    gall -= gall.min(0)
    for i, gc in zip(idx, grid):
        members = np.where((gall == gc).all(1))[0]
        assert i == members.min()
    prio = rng.permutation(500)
    idx2, _ = G.grid_sample(coord, 0.01, prio=prio)
    for i, gc in zip(idx2, grid):
        members = np.where((gall == gc).all(1))[0]
        assert i == members[np.argmin(prio[members])]
