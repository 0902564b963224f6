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
