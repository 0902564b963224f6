"""Host-side (device-agnostic) parts of the PRODUCT modules, checked on the CPU without launching a kernel:
constructor / state_dict compatibility with fixtures saved from the reference classes, the DDPM schedule and
sampler coefficients against the oracle, no-CPU-fallback behaviour."""
import numpy as np
import pytest
import torch

from tests._golden_act import GOLDEN_ACT
from tests._golden_act import load as load_act
from tests._golden_dp import GOLDEN_DP, GOLDEN_DPENC, encoder_kwargs, load, load_encoder


@pytest.mark.parametrize("path", GOLDEN_ACT)
def test_act_product_state_dict_is_reference_compatible(path):
    from pointcloudmatters_b200.act import build_policy

    cfg, state, *_rest, rlbench = load_act(path)
    model = build_policy(cfg, rlbench)
    assert sorted(model.state_dict().keys()) == sorted(state.keys())
    model.load_state_dict(state)  # shapes


@pytest.mark.parametrize("path", GOLDEN_DP)
def test_dp_product_state_dict_is_reference_compatible(path):
    from pointcloudmatters_b200.diffusion import build_dp_policy

    cfg, state, *_ = load(path)
    model = build_dp_policy(cfg)
    assert sorted(model.state_dict().keys()) == sorted(k for k in state if not k.startswith("normalizer."))
    model.load_state_dict(state)
    assert sorted(model.state_dict().keys()) == sorted(state.keys())
    assert model.model.cond_dim == cfg["diffusion_step_embed_dim"] + 2 * (cfg["projector_channels"][-1] + cfg["qpos_dim"]) + cfg["goal_dim"]


@pytest.mark.parametrize("path", GOLDEN_DPENC)
def test_dp_encoder_variants_state_dict_is_reference_compatible(path):
    from pointcloudmatters_b200.diffusion import PCDObsEncoder
    from pointcloudmatters_b200.pointnet import PointNet

    cfg, state, *_ = load_encoder(path)
    sm, kw = encoder_kwargs(cfg)
    enc = PCDObsEncoder(sm, PointNet(6, cfg["backbone_classes"]), **kw)
    assert sorted(enc.state_dict().keys()) == sorted(state.keys())
    enc.load_state_dict(state)


def test_product_ddpm_scheduler_equals_oracle():
    from oracle.dp_oracle import DDPMSchedule
    from pointcloudmatters_b200.diffusion import DDPMScheduler

    a, b = DDPMScheduler(num_train_timesteps=100), DDPMSchedule(num_train_timesteps=100)
    assert torch.equal(a.alphas_cumprod, b.alphas_cumprod)
    g = torch.Generator().manual_seed(0)
    x, e, n = (torch.randn(4, 16, 7, generator=g) for _ in range(3))
    t = torch.tensor([0, 17, 50, 99])
    assert torch.equal(a.add_noise(x, e, t), b.add_noise(x, e, t))
    for steps in (100, 10, 7):
        a.set_timesteps(steps); b.set_timesteps(steps)
        assert torch.equal(a.timesteps, b.timesteps) and a.timesteps[-1] == 0
        for ti in a.timesteps.tolist():
            c = a.step_coefficients(ti)
            x0 = ((x - c[0] * e) * c[1]).clamp(-1, 1)
            mine = c[2] * x0 + c[3] * x + c[4] * n
            assert torch.allclose(mine, b.step(e, ti, x, noise=n).prev_sample, rtol=1e-5, atol=1e-5), (steps, ti)
        assert float(a.step_coefficients(0)[4]) == 0.0  # no noise at the last step


def test_product_ops_refuse_cpu_tensors():
    """No CPU fallback: the operator layer raises instead of silently computing on the host."""
    from pointcloudmatters_b200 import functional as PF
    from pointcloudmatters_b200 import functional_unet as UF
    from pointcloudmatters_b200._lib import PcmError

    x = torch.randn(4, 16)
    with pytest.raises(PcmError):
        PF.linear(x, torch.randn(8, 16))
    with pytest.raises(PcmError):
        UF.conv1d_cl(torch.randn(2, 4, 8), torch.randn(8, 8, 3))
    with pytest.raises(PcmError):
        UF.mish(x)
    with pytest.raises(PcmError):
        PF.feed_forward(torch.randn(4, 32), torch.nn.Linear(32, 16), torch.nn.Linear(16, 32), 0.1, True)
