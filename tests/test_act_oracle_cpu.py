"""Pins oracle/act_oracle.py (the CPU restatement of ACTPCD / ACTRLBenchPCD / Transformer) against
fixtures produced by the REFERENCE's own modules (oracle/gen_golden_act.py): same state_dict keys,
same outputs, same gradients, same BatchNorm running statistics, fp32 tolerance 2e-5."""
import numpy as np
import pytest
import torch

from oracle.act_oracle import build_oracle_policy
from tests._golden_act import GOLDEN_ACT, grad_summary, load

RTOL, ATOL = 2e-4, 2e-5


@pytest.mark.parametrize("path", GOLDEN_ACT)
def test_oracle_policy_matches_reference_modules(path):
    cfg, state, batch, out, grads, post, nograd, rlbench = load(path)
    model = build_oracle_policy(cfg, rlbench)
    # identical state_dict surface (keys AND shapes) as the reference module -- checkpoint contract
    ref_keys = {k: tuple(v.shape) for k, v in state.items()}
    own_keys = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert ref_keys == own_keys
    model.load_state_dict(state)
    model.train()
    d = model(batch)
    for k in ("a_hat", "is_pad_hat", "mu", "logvar", "loss", "action_loss", "kl_loss"):
        np.testing.assert_allclose(d[k].detach().numpy(), out[k], rtol=RTOL, atol=ATOL, err_msg=k)
    d["loss"].backward()
    got_nograd = sorted(k for k, p in model.named_parameters() if p.grad is None)
    assert got_nograd == sorted(nograd)  # is_pad_head.* never receives a gradient
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        want = grads[k]
        got = grad_summary(p.grad)
        scale = max(want[0], 1e-6)
        np.testing.assert_allclose(got / scale, want / scale, rtol=1e-3, atol=2e-4, err_msg=k)
    # dead decoder layers: exactly-zero gradients (only [0] of the intermediate stack is consumed)
    for k, p in model.named_parameters():
        if k.startswith("transformer.decoder.layers.") and not k.startswith("transformer.decoder.layers.0."):
            assert float(p.grad.abs().max()) == 0.0, k
    sd = model.state_dict()
    for k, v in post.items():
        np.testing.assert_allclose(sd[k].numpy(), v, rtol=RTOL, atol=ATOL, err_msg=k)


@pytest.mark.parametrize("path", [p for p in GOLDEN_ACT if p.endswith(("_small.npz", "_h128.npz"))])
def test_oracle_eval_mode_matches_reference_fixture(path):
    """Inference branch (no actions: latent 0, BatchNorm on running statistics; RLBench: rot6d -> quaternion) against
    the reference module's own eval-mode output stored in the fixture (`eval/a_hat`)."""
    import numpy as np

    from oracle.act_oracle import build_oracle_policy

    cfg, state, batch, _out, _g, _p, _n, rlbench = load(path)
    want = np.load(path)["eval/a_hat"]
    model = build_oracle_policy(cfg, rlbench).eval()
    model.load_state_dict(state)
    with torch.no_grad():
        got = model({k: v for k, v in batch.items() if k in ("pcds", "qpos", "goal_cond")})["a_hat"].numpy()
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5)
