"""The Diffusion-Policy oracle (oracle/dp_oracle.py) against fixtures produced by the REFERENCE's own
modules (oracle/gen_golden_dp.py): identical state_dict keys, loss and gradients in fp32 on the CPU."""
import numpy as np
import pytest
import torch

from tests._golden_act import grad_summary
from tests._golden_dp import GOLDEN_DP, GOLDEN_DPENC, encoder_kwargs, load, load_encoder, load_prediction


def test_fixtures_present():
    assert len(GOLDEN_DP) >= 2


@pytest.mark.parametrize("path", GOLDEN_DP)
def test_oracle_matches_reference_fixture(path):
    from oracle.dp_oracle import build_oracle_dp

    cfg, state, batch, loss, grads, post, nograd = load(path)
    model = build_oracle_dp(cfg).train()
    # drop-in checkpoint compatibility (the normaliser's fields appear when it is fit / loaded)
    assert sorted(model.state_dict().keys()) == sorted(k for k in state if not k.startswith("normalizer."))
    model.load_state_dict(state)
    assert sorted(model.state_dict().keys()) == sorted(state.keys())
    out = model.compute_loss(batch)
    assert abs(float(out["loss"]) - loss) <= 1e-5 * abs(loss)
    out["loss"].backward()
    assert sorted(k for k, p in model.named_parameters() if p.requires_grad and p.grad is None) == sorted(nograd)
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        want, got = grads[k], grad_summary(p.grad)
        np.testing.assert_allclose(got, want, rtol=2e-3, atol=2e-6 + 1e-4 * abs(want[0]), err_msg=k)
    sd = model.state_dict()
    for k, v in post.items():
        np.testing.assert_allclose(sd[k].numpy(), v, rtol=1e-5, atol=1e-6, err_msg=k)


@pytest.mark.parametrize("path", GOLDEN_DP)
def test_oracle_predict_action_matches_reference_fixture(path):
    """Sampling loop (normalise -> encode -> 10 x [denoiser, DDPM step] -> unnormalise -> slice) against the
    reference's own `predict_action` run with the same recorded noise draws."""
    from oracle.dp_oracle import build_oracle_dp

    cfg, state, batch, *_ = load(path)
    noises, action, action_pred = load_prediction(path)
    model = build_oracle_dp(dict(cfg, num_inference_steps=10)).eval()
    model.load_state_dict(state)
    obs = {k: v for k, v in batch.items() if k in ("obs", "goal")}
    out = model.predict_action(obs, noises=noises)
    assert torch.allclose(out["action_pred"], action_pred, rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["action"], action, rtol=1e-4, atol=1e-5)
    assert out["action"].shape == (action_pred.shape[0], 8, cfg["action_dim"])


@pytest.mark.parametrize("path", GOLDEN_DPENC)
def test_oracle_encoder_variants_match_reference_fixture(path):
    """`use_mask` (+ bg_ratio) and `pre_sample` variants of PCDObsEncoder (pcd_obs_encoder.py:91-93,133-177,201-218)."""
    from oracle.act_oracle import OraclePointNet
    from oracle.dp_oracle import OraclePCDObsEncoder

    cfg, state, obs, feats, probe, grads, post = load_encoder(path)
    sm, kw = encoder_kwargs(cfg)
    enc = OraclePCDObsEncoder(sm, OraclePointNet(6, cfg["backbone_classes"]), **kw).train()
    assert sorted(enc.state_dict().keys()) == sorted(state.keys())
    enc.load_state_dict(state)
    out = enc(obs)
    assert torch.allclose(out, feats, rtol=1e-4, atol=1e-5)
    (out * probe).sum().backward()
    for k, p in enc.named_parameters():
        if p.grad is not None:
            np.testing.assert_allclose(grad_summary(p.grad), grads[k], rtol=2e-3, atol=2e-6 + 1e-4 * abs(grads[k][0]), err_msg=k)
    for k, v in post.items():
        np.testing.assert_allclose(enc.state_dict()[k].numpy(), v, rtol=1e-5, atol=1e-6, err_msg=k)


def test_ddpm_schedule_known_values():
    """squaredcos_cap_v2 (Nichol & Dhariwal): closed-form checks of the restated scheduler."""
    import math

    from oracle.dp_oracle import DDPMSchedule

    s = DDPMSchedule(num_train_timesteps=100)
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    assert abs(float(s.alphas_cumprod[0]) - ab(0.01) / ab(0.0)) < 1e-6
    # the cap (beta <= 0.999) only binds at the last step: acp[98] is the uncapped closed form
    assert abs(float(s.alphas_cumprod[98]) - ab(0.99) / ab(0.0)) < 1e-6
    assert float(s.alphas_cumprod[99]) > 0.0 and torch.all(s.alphas_cumprod[1:] < s.alphas_cumprod[:-1])
    x, n = torch.ones(2, 3, 4), torch.full((2, 3, 4), 2.0)
    t = torch.tensor([0, 50])
    y = s.add_noise(x, n, t)
    a = s.alphas_cumprod[t]
    assert torch.allclose(y[:, 0, 0], a.sqrt() + 2 * (1 - a).sqrt())
