"""Generate golden vectors from the UNMODIFIED reference pointops kernels (run on a B200):

    python tests/golden/gen_golden_ref_gpu.py gpurun_out/golden

Writes ref_pointops_<case>.npz (inputs + FPS / kNN / ball / random-ball outputs of
oracle/_ref/libpointops_ref.so).  The small files are then committed under tests/golden/ and
pin the CPU oracle in tests/test_oracle_cpu.py::test_oracle_matches_reference_golden.
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import _ref  # noqa: E402
from tests._data import clouds  # noqa: E402

GOLDEN_CASES = [
    # name, b, n, m, kind, ragged, nsample, max_r, min_r
    ("uniform_b2_n512", 2, 512, 256, "uniform", False, 16, 0.15, 0.0),
    ("ragged_b3_n1000", 3, 1000, 300, "uniform", True, 16, 0.2, 0.05),
    ("lattice_b2_n700", 2, 700, 128, "lattice", True, 8, 0.3, 0.0),
    ("dup_b2_n600", 2, 600, 200, "dup", False, 12, 0.12, 0.0),
    ("tiny_b4_n37", 4, 37, 16, "uniform", True, 16, 0.5, 0.0),
    ("big_b1_n3000", 1, 3000, 512, "uniform", False, 32, 0.25, 0.0),
]


def main(out_dir):
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    dev = torch.device("cuda:0")
    for name, b, n, m, kind, ragged, k, rmax, rmin in GOLDEN_CASES:
        xyz, off, noff = clouds(b, n, m, seed=1234, kind=kind, ragged=ragged)
        rng = np.random.default_rng(7)
        starts = np.concatenate([[0], off[:-1]])
        order = np.concatenate([s + rng.permutation(e - s) for s, e in zip(starts, off)]).astype(np.int32)
        t_xyz, t_off, t_noff = torch.from_numpy(xyz).to(dev), torch.from_numpy(off).to(dev), torch.from_numpy(noff).to(dev)
        fps = _ref.farthest_point_sampling(t_xyz, t_off, t_noff)
        q = t_xyz[fps.long()].contiguous()
        ki, kd = _ref.knn_query(k, t_xyz, t_off, q, t_noff)
        bi, bd = _ref.ball_query(k, rmax, rmin, t_xyz, t_off, q, t_noff)
        ri, rd = _ref.random_ball_query(k, rmax, rmin, torch.from_numpy(order).to(dev), t_xyz, t_off, q, t_noff)
        np.savez_compressed(out / f"ref_pointops_{name}.npz", xyz=xyz, offset=off, new_offset=noff, order=order,
                            nsample=k, max_radius=rmax, min_radius=rmin, fps_idx=fps.cpu().numpy(),
                            knn_idx=ki.cpu().numpy(), knn_dist2=kd.cpu().numpy(), ball_idx=bi.cpu().numpy(),
                            ball_dist2=bd.cpu().numpy(), rball_idx=ri.cpu().numpy(), rball_dist2=rd.cpu().numpy())
        print("wrote", name)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
