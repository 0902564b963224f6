"""CPU tests of the oracle (oracle/pointops_oracle.c) against independent numpy restatements,
the reference's documented conventions, and (when present) golden vectors produced by the
UNMODIFIED reference kernels on a B200 (tests/golden/ref_pointops_*.npz, made by
tests/golden/gen_golden_ref_gpu.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import pointops_oracle as O
from tests._data import CASES, clouds


def _d2(q, p):
    """float32 distance with the reference contraction, emulated via float64 (exact products)."""
    dx = (q[..., 0] - p[..., 0]).astype(np.float32)
    dy = (q[..., 1] - p[..., 1]).astype(np.float32)
    dz = (q[..., 2] - p[..., 2]).astype(np.float32)
    t = (dy.astype(np.float64) * dy).astype(np.float32)  # FMUL lands on the y term (reference SASS)
    t = (dx.astype(np.float64) * dx + t).astype(np.float32)
    return (dz.astype(np.float64) * dz + t).astype(np.float32)


def test_opt_n_threads_rule():
    # cuda_utils.h:11-14
    for n, want in [(1, 1), (2, 2), (3, 2), (8, 8), (37, 32), (512, 512), (1000, 512), (1024, 1024), (4096, 1024)]:
        assert O.opt_n_threads(n) == want


@pytest.mark.parametrize("case", CASES)
def test_fps_greedy_property(case):
    b, n, m, kind, ragged = case
    xyz, off, noff = clouds(b, n, m, seed=11, kind=kind, ragged=ragged)
    idx = O.farthest_point_sampling(xyz, off, noff)
    starts = np.concatenate([[0], off[:-1]])
    nstarts = np.concatenate([[0], noff[:-1]])
    for c in range(b):
        s, e, ns, ne = starts[c], off[c], nstarts[c], noff[c]
        sel = idx[ns:ne]
        assert sel[0] == s  # first pick = first point of the cloud (sampling_cuda_kernel.cu:39)
        assert ((sel >= s) & (sel < e)).all()
        mind = np.full(e - s, 1e10, dtype=np.float32)
        for j in range(1, len(sel)):
            mind = np.minimum(mind, _d2(xyz[s:e], xyz[sel[j - 1]][None]))
            # the pick maximises the running min distance (ties allowed)
            assert mind[sel[j] - s] == mind.max()
        if kind == "uniform" and ne - ns <= e - s:
            assert len(set(sel.tolist())) == len(sel)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("k", [1, 3, 16])
def test_knn_sorted_and_matches_bruteforce(case, k):
    b, n, m, kind, ragged = case
    xyz, off, noff = clouds(b, n, m, seed=5, kind=kind, ragged=ragged)
    fidx = O.farthest_point_sampling(xyz, off, noff)
    q = xyz[fidx]
    idx, d2 = O.knn_query(k, xyz, off, q, noff, squared=True)
    starts = np.concatenate([[0], off[:-1]])
    nstarts = np.concatenate([[0], noff[:-1]])
    for c in range(b):
        s, e = starts[c], off[c]
        for qi in range(nstarts[c], noff[c]):
            d = _d2(q[qi][None], xyz[s:e])
            kk = min(k, e - s)
            assert (np.diff(d2[qi, :kk]) >= 0).all()  # ascending (heap_sort)
            assert np.array_equal(np.sort(d)[:kk], d2[qi, :kk])  # same multiset of distances
            assert np.array_equal(d[idx[qi, :kk] - s], d2[qi, :kk])  # indices consistent
            assert (idx[qi, kk:] == -1).all() and (d2[qi, kk:] == np.float32(1e10)).all()
            if kind == "uniform":
                assert np.array_equal(np.argsort(d, kind="stable")[:kk] + s, idx[qi, :kk])


def test_knn_returns_sqrt_distance_and_self_query():
    xyz, off, _ = clouds(2, 64, None, seed=3)
    idx, dist = O.knn_query(4, xyz, off)
    assert (idx[:, 0] == np.arange(xyz.shape[0])).all() and (dist[:, 0] == 0).all()
    _, d2 = O.knn_query(4, xyz, off, squared=True)
    assert np.array_equal(dist, np.sqrt(d2))


def test_ball_query_conventions():
    xyz, off, noff = clouds(3, 400, 50, seed=9, ragged=True)
    q = xyz[O.farthest_point_sampling(xyz, off, noff)]
    ns, rmax, rmin = 8, 0.12, 0.02
    idx, d2 = O.ball_query(ns, rmax, rmin, xyz, off, q, noff, squared=True)
    starts = np.concatenate([[0], off[:-1]])
    nstarts = np.concatenate([[0], noff[:-1]])
    for c in range(3):
        s, e = starts[c], off[c]
        for qi in range(nstarts[c], noff[c]):
            d = _d2(q[qi][None], xyz[s:e])
            hit = (d.astype(np.float64) <= 1e-5) | ((d >= np.float32(rmin) ** 2) & (d < np.float32(rmax) ** 2))
            cnt = int(hit.sum())
            if cnt <= ns:
                assert set(idx[qi, :cnt].tolist()) == set((np.nonzero(hit)[0] + s).tolist())
                assert (idx[qi, cnt:] == -1).all() and (d2[qi, cnt:] == np.float32(1e10)).all()
            else:
                # strided-subsample branch: every index is a hit, dist2 holds the INDEX (sic, .cu:120)
                assert all(hit[i - s] for i in idx[qi])
                assert np.array_equal(d2[qi], idx[qi].astype(np.float32))


def test_random_ball_query_first_hits_in_order():
    xyz, off, noff = clouds(2, 300, 40, seed=2)
    q = xyz[O.farthest_point_sampling(xyz, off, noff)]
    rng = np.random.default_rng(0)
    order = np.concatenate([rng.permutation(300), 300 + rng.permutation(300)]).astype(np.int32)
    idx, d2 = O.random_ball_query(6, 0.15, 0.0, xyz, off, q, noff, order, squared=True)
    for qi in range(q.shape[0]):
        c = qi // 40
        seq = order[c * 300:(c + 1) * 300]
        d = _d2(q[qi][None], xyz[seq])
        hits = seq[(d.astype(np.float64) <= 1e-5) | (d < np.float32(0.15) ** 2)][:6]
        assert np.array_equal(idx[qi, :len(hits)], hits)
        assert (idx[qi, len(hits):] == -1).all()


def test_gather_scatter_ops_against_numpy():
    rng = np.random.default_rng(4)
    n, m, ns, c = 50, 20, 5, 7
    inp = rng.standard_normal((n, c)).astype(np.float32)
    idx = rng.integers(0, n, (m, ns)).astype(np.int32)
    assert np.array_equal(O.grouping_forward(inp, idx), inp[idx])
    go = rng.standard_normal((m, ns, c)).astype(np.float32)
    ref = np.zeros((n, c)); np.add.at(ref, idx.reshape(-1), go.reshape(-1, c).astype(np.float64))
    np.testing.assert_allclose(O.grouping_backward(go, idx, n), ref, rtol=1e-5, atol=1e-5)
    w = rng.random((m, ns)).astype(np.float32)
    np.testing.assert_allclose(O.interpolation_forward(inp, idx, w), (inp[idx] * w[..., None]).sum(1), rtol=1e-5, atol=1e-5)
    i1 = rng.standard_normal((n, c)).astype(np.float32)
    idx2 = rng.integers(0, n, (n, ns)).astype(np.int32)
    assert np.array_equal(O.subtraction_forward(i1, inp, idx2), i1[:, None, :] - inp[idx2])
    pos = rng.standard_normal((n, ns, c)).astype(np.float32)
    wt = rng.standard_normal((n, ns, 1)).astype(np.float32)
    np.testing.assert_allclose(O.aggregation_forward(inp, pos, wt, idx2), ((inp[idx2] + pos) * wt).sum(1), rtol=1e-4, atol=1e-5)


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_pointops_*.npz")))


@pytest.mark.skipif(not GOLDEN, reason="golden vectors from the reference kernels not generated yet")
@pytest.mark.parametrize("path", GOLDEN)
def test_oracle_matches_reference_golden(path):
    """Pins the oracle: bit-exact against outputs of the unmodified reference CUDA kernels."""
    g = np.load(path)
    xyz, off, noff = g["xyz"], g["offset"], g["new_offset"]
    fps = O.farthest_point_sampling(xyz, off, noff)
    assert np.array_equal(fps, g["fps_idx"])
    q = xyz[g["fps_idx"]]
    k = int(g["nsample"])
    ki, kd2 = O.knn_query(k, xyz, off, q, noff, squared=True)
    assert np.array_equal(ki, g["knn_idx"]) and np.array_equal(kd2, g["knn_dist2"])
    bi, bd2 = O.ball_query(k, float(g["max_radius"]), float(g["min_radius"]), xyz, off, q, noff, squared=True)
    assert np.array_equal(bi, g["ball_idx"]) and np.array_equal(bd2, g["ball_dist2"])
    ri, rd2 = O.random_ball_query(k, float(g["max_radius"]), float(g["min_radius"]), xyz, off, q, noff, g["order"], squared=True)
    assert np.array_equal(ri, g["rball_idx"]) and np.array_equal(rd2, g["rball_dist2"])
