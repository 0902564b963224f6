"""Host-side rollout helpers against golden vectors produced by the reference's own class source
(oracle/gen_golden_misc.py executes `class TemporalAgg` cut out of src/utils/misc.py:88-140)."""
import os

import numpy as np


def test_temporal_agg_matches_reference_vectors():
    from pointcloudmatters_b200.inference import TemporalAgg

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "temporal_agg.npz"))
    for name in ("a", "b"):
        chunk, dim, k, steps = g[f"{name}/cfg"]
        agg = TemporalAgg(apply=True, action_dim=int(dim), chunk_size=int(chunk), k=float(k))
        got = np.stack([agg(c) for c in g[f"{name}/chunks"]])
        np.testing.assert_array_equal(got, g[f"{name}/actions"])  # same numpy arithmetic: bit-exact
    passthrough = TemporalAgg(apply=False)
    assert passthrough(np.arange(6).reshape(2, 3))[1] == 1


def _golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "misc_rotation_normalizer.npz"))


def test_rot6d_to_quaternion_matches_reference_vectors():
    """The rot6d -> matrix -> quaternion conversion of ACTRLBenchPCD's inference branch (act.py:785-795) against
    vectors from the reference's rotation_conversions.py (incl. the w = 0 / 180-degree branches)."""
    import torch

    from pointcloudmatters_b200.act import _matrix_to_quaternion, _rotation_6d_to_matrix

    g = _golden()
    d6 = torch.from_numpy(g["rot/d6"])
    m = _rotation_6d_to_matrix(d6)
    np.testing.assert_allclose(m.numpy(), g["rot/matrix"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(_matrix_to_quaternion(m).numpy(), g["rot/quat"], rtol=1e-5, atol=1e-6)


def test_linear_normalizer_fit_matches_reference_vectors():
    """`LinearNormalizer.fit` / normalize / unnormalize (normalizer.py:195-300) in limits, gaussian and
    fit_offset=False modes, with a constant channel (range_eps branch), against the reference's own class."""
    import torch

    from pointcloudmatters_b200.diffusion import LinearNormalizer

    g = _golden()
    data, x = torch.from_numpy(g["norm/data"]), torch.from_numpy(g["norm/x"])
    for mode, kw in (("limits", {}), ("gaussian", {}), ("limits_nooffset", {"fit_offset": False})):
        n = LinearNormalizer().fit({"action": data}, mode=mode.split("_")[0], **kw)
        np.testing.assert_allclose(n.params_dict["action"]["scale"].numpy(), g[f"norm/{mode}/scale"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(n.params_dict["action"]["offset"].numpy(), g[f"norm/{mode}/offset"], rtol=1e-6, atol=1e-6)
        y = n.normalize_field("action", x)
        np.testing.assert_allclose(y.numpy(), g[f"norm/{mode}/y"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(n.normalize_field("action", y, forward=False).numpy(), g[f"norm/{mode}/back"], rtol=1e-5, atol=1e-5)
    # state_dict round trip through the reference's key layout
    m = LinearNormalizer()
    m.load_state_dict(n.state_dict())
    assert sorted(m.state_dict().keys()) == sorted(n.state_dict().keys())
    assert "params_dict.action.input_stats.min" in m.state_dict()
