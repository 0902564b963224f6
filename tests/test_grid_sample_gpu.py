"""GPU voxel-grid subsampling + collation (csrc/grid_sample.cu, pointcloudmatters_b200/data_gpu.py) -- integer work,
bit-exact: against the numpy oracle on every output (selected rows included) and against the REFERENCE classes' own
outputs for everything that does not depend on their unstable argsort (tests/golden/grid_sample_ref.npz)."""
import numpy as np
import pytest
import torch

from oracle import grid_sample_oracle as G
from tests.test_grid_sample_cpu import CASES, GOLD, load_case

pytestmark = pytest.mark.gpu


def _run(clouds, gs, **kw):
    from pointcloudmatters_b200.data_gpu import collate_raw_clouds

    out = collate_raw_clouds(clouds, "cuda", grid_size=gs, return_index=True, **kw)
    torch.cuda.synchronize()
    return out


def _local_index(out, clouds):
    starts = np.concatenate([[0], np.cumsum([c.shape[0] for c, _ in clouds])[:-1]])
    idx = out["index"].cpu().numpy()
    new_off = out["offset"].cpu().numpy()
    res, s = [], 0
    for st, e in zip(starts, new_off):
        res.append(idx[s:e] - st)
        s = e
    return np.concatenate(res)


@pytest.mark.parametrize("case", CASES)
def test_matches_oracle_and_reference_fixture(case):
    g = np.load(GOLD)
    clouds, gs = load_case(g, case)
    out = _run(clouds, gs)
    want = G.grid_sample_collate(clouds, gs)
    assert np.array_equal(out["offset"].cpu().numpy(), want["offset"])
    assert np.array_equal(out["grid_coord"].cpu().numpy(), want["grid_coord"])
    assert np.array_equal(_local_index(out, clouds), want["index"])
    assert np.array_equal(out["coord"].cpu().numpy(), want["coord"])
    assert np.array_equal(out["feat"].cpu().numpy(), want["feat"])
    assert out["n_max"] == int(np.diff(np.concatenate([[0], want["offset"]])).max())
    # the reference classes' own outputs (tie-order independent parts)
    assert np.array_equal(out["offset"].cpu().numpy(), g[f"{case}/out_offset"])
    assert np.array_equal(out["grid_coord"].cpu().numpy(), g[f"{case}/out_grid_coord"])


def test_large_cloud_takes_the_global_sort_path_and_random_priorities():
    rng = np.random.default_rng(5)
    clouds = [((rng.uniform(-0.5, 0.5, (40000, 3))).astype(np.float32), rng.integers(0, 256, (40000, 3)).astype(np.float32)),
              ((rng.uniform(-0.1, 0.1, (300, 3))).astype(np.float32), rng.integers(0, 256, (300, 3)).astype(np.float32)),
              (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)),
              ((rng.uniform(-0.2, 0.2, (9000, 3))).astype(np.float32), rng.integers(0, 256, (9000, 3)).astype(np.float32))]
    want = G.grid_sample_collate(clouds, 0.005)
    assert np.diff(np.concatenate([[0], want["offset"]])).max() > 8192  # > SORT_CAP voxels in one cloud
    out = _run(clouds, 0.005)
    for k in ("offset", "grid_coord", "coord", "feat"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k
    # explicit priorities (the train-mode mechanism): survivor = argmin priority inside the voxel
    prios = [rng.permutation(c.shape[0]) for c, _ in clouds]
    want = G.grid_sample_collate(clouds, 0.02, prios=prios)
    out = _run(clouds, 0.02, prio=torch.from_numpy(np.concatenate(prios)).cuda())
    for k in ("offset", "grid_coord", "coord", "feat"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k
    assert np.array_equal(_local_index(out, clouds), want["index"])
    # float32 division (numpy 1.x promotion) variant
    want = G.grid_sample_collate(clouds, 0.005, f32_div=True)
    out = _run(clouds, 0.005, f32_div=True)
    assert np.array_equal(out["grid_coord"].cpu().numpy(), want["grid_coord"])


def test_train_mode_draws_a_member_of_every_voxel_and_feeds_the_policy():
    """mode='train': same voxels / order as test mode, some other member; the result has the batch contract's dtypes and
    drives one ACT training step with the n_max hint (sync-free FPS)."""
    from pointcloudmatters_b200.act import build_policy
    from pointcloudmatters_b200.bc_module import ACTBCModule
    from pointcloudmatters_b200.data_gpu import collate_raw_clouds

    rng = np.random.default_rng(9)
    clouds = [((rng.normal(0, 0.03, (4000, 3))).astype(np.float32), rng.integers(0, 256, (4000, 3)).astype(np.float32))
              for _ in range(4)]
    a = collate_raw_clouds(clouds, "cuda", grid_size=0.005, mode="test", return_index=True)
    b = collate_raw_clouds(clouds, "cuda", grid_size=0.005, mode="train", seed=3, return_index=True)
    assert torch.equal(a["grid_coord"], b["grid_coord"]) and torch.equal(a["offset"], b["offset"])
    assert not torch.equal(a["index"], b["index"])
    gs = torch.floor(b["coord"].double() / 0.005)
    # every survivor lies in the voxel it represents: same grid cell as test mode's survivor
    assert torch.equal(gs, torch.floor(a["coord"].double() / 0.005))
    assert b["coord"].dtype == torch.float32 and b["grid_coord"].dtype == torch.int64 and b["feat"].shape[1] == 6
    cfg = dict(hidden_dim=128, nhead=2, dim_feedforward=32, enc_layers=1, dec_layers=1, dropout=0.0, num_queries=12,
               action_dim=7, qpos_dim=9, goal_cond_dim=3, latent_dim=32, kl_weight=10.0, pcd_npoints=64, pcd_nsample=16)
    torch.manual_seed(0)
    module = ACTBCModule(build_policy(cfg).cuda().train(), total_steps=100)
    batch = {"pcds": {k: b[k] for k in ("coord", "grid_coord", "feat", "offset", "n_max")}, "qpos": torch.randn(4, 9).cuda(),
             "actions": torch.randn(4, 12, 7).cuda(), "is_pad": torch.zeros(4, 12, dtype=torch.bool).cuda(),
             "goal_cond": torch.randn(4, 3).cuda()}
    loss = float(module.training_step(batch, 0))
    assert loss == loss
