"""Parity at the REAL dimensions of the BASELINE configs: hidden 512, 8 heads, 4 encoder +
7 decoder layers -- cfg-2 (N=1024, M=512, S=515) and cfg-4 (RLBench head, N=4096, M=2048, S=2051) -- against the
oracle port (oracle/act_oracle.py, pinned to the reference modules by tests/test_act_oracle_cpu.py) on the same
weights and inputs: outputs, the three losses and a per-tensor gradient summary, dropout 0 and injected
reparametrisation noise.  Small batches (4 / 2 samples) keep the fp32 CPU oracle to seconds; every kernel runs
at its production tile / sequence shape because those depend on N, M, S and the width, not on the batch.

Tolerances as tests/test_act_gpu.py (bf16 tensor-core operands vs a pure fp32 oracle): outputs rel-L2 <= 3e-2,
losses <= 2e-2, gradients <= 6e-2 per tensor on the scale max(|g|, 1e-3 * largest gradient norm); the dead
decoder layers 1..6 must be exactly zero."""
import numpy as np
import pytest
import torch

from tests._golden_act import grad_summary

pytestmark = pytest.mark.gpu

OUT_TOL, LOSS_TOL, GRAD_TOL = 3e-2, 2e-2, 6e-2


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


def _run(cfg, rlbench, batch_size, n_points, seed):
    from oracle.act_oracle import build_oracle_policy
    from pointcloudmatters_b200._lib import lib
    from pointcloudmatters_b200.act import build_policy
    from pointcloudmatters_b200.data import synthetic_act_batch, to_device

    torch.manual_seed(seed)
    policy = build_policy(cfg, rlbench=rlbench).cuda().train()
    with torch.no_grad():  # de-trivialise the parameters the default init leaves at 0 / 1
        for pn, p in policy.named_parameters():
            if pn.endswith("bias") or "norm" in pn or pn.endswith("bn.weight"):
                p.add_(0.1 * torch.randn_like(p))
    oracle = build_oracle_policy(cfg, rlbench=rlbench).train()
    oracle.load_state_dict({k: v.detach().cpu() for k, v in policy.state_dict().items()})
    batch = synthetic_act_batch(batch_size, n_points, num_queries=cfg["num_queries"], action_dim=cfg["action_dim"],
                                qpos_dim=cfg["qpos_dim"], goal_cond_dim=cfg["goal_cond_dim"], seed=seed, ragged=True)
    if rlbench:
        batch["actions"][..., -2:] = torch.rand(batch_size, cfg["num_queries"], 2)
    eps = torch.randn(batch_size, cfg["latent_dim"])
    ob = {k: (dict(v) if isinstance(v, dict) else v) for k, v in batch.items()}
    ob["pcds"].pop("n_max")
    ob["_eps"] = eps
    want = oracle(ob)
    want["loss"].backward()
    gb = to_device(batch, "cuda")
    gb["pcds"]["n_max"] = batch["pcds"]["n_max"]
    gb["_eps"] = eps.cuda()
    before = dict(lib.calls)
    got = policy(gb)
    got["loss"].backward()
    torch.cuda.synchronize()
    ran = {k: v - before.get(k, 0) for k, v in lib.calls.items()}
    n_mha = 2 * cfg["enc_layers"] + 2 * cfg["dec_layers"]
    assert ran.get("pcm_flash_attn_fwd", 0) >= n_mha and ran.get("pcm_flash_attn_bwd", 0) >= n_mha, ran
    for k in ("a_hat", "mu", "logvar"):
        assert _rel(got[k].detach().cpu(), want[k].detach()) <= OUT_TOL, k
    for k in ("loss", "action_loss", "kl_loss"):
        assert abs(float(got[k]) - float(want[k])) <= LOSS_TOL * abs(float(want[k])) + 1e-5, k
    og = dict(oracle.named_parameters())
    floor = 1e-3 * max(float(p.grad.norm()) for p in og.values() if p.grad is not None)
    bad = {}
    for k, p in policy.named_parameters():
        if og[k].grad is None:
            assert p.grad is None, k
            continue
        if k.startswith("transformer.decoder.layers.") and not k.startswith("transformer.decoder.layers.0."):
            assert float(p.grad.abs().max()) == 0.0 and float(og[k].grad.abs().max()) == 0.0, k
            continue
        w, g = grad_summary(og[k].grad), grad_summary(p.grad)
        scale = max(w[0], floor)
        err = max(abs(g[0] - w[0]) / scale, np.abs(g[2:] - w[2:]).max() / scale)
        if err > GRAD_TOL:
            bad[k] = err
    assert not bad, bad
    return policy, gb, float(want["loss"])


def test_cfg2_full_dims_match_oracle():
    """hidden 512 / 8 heads / 4 enc + 7 dec / N=1024 / M=512 / 100 queries (bench.py's CFG2), B=4."""
    cfg = dict(hidden_dim=512, nhead=8, dim_feedforward=32, enc_layers=4, dec_layers=7, dropout=0.0, num_queries=100,
               action_dim=7, qpos_dim=9, goal_cond_dim=3, latent_dim=32, kl_weight=10.0, pcd_npoints=512, pcd_nsample=16)
    _run(cfg, False, 4, 1024, seed=21)


def test_cfg4_full_dims_match_oracle():
    """RLBench cfg-4 with ALL layers: N=4096, M=2048 (S=2051), action_dim 11 (rot6d + gripper + collision),
    512-d goal embedding, B=2 -- then one full optimizer step on the same batch."""
    from pointcloudmatters_b200.bc_module import ACTBCModule

    cfg = dict(hidden_dim=512, nhead=8, dim_feedforward=32, enc_layers=4, dec_layers=7, dropout=0.0, num_queries=100,
               action_dim=11, qpos_dim=4, goal_cond_dim=512, latent_dim=32, kl_weight=10.0, pcd_npoints=2048,
               pcd_nsample=16, collision=True, position_loss_weight=3.0)
    policy, gb, want_loss = _run(cfg, True, 2, 4096, seed=22)
    policy.zero_grad(set_to_none=True)
    module = ACTBCModule(policy, total_steps=100)
    l0 = float(module.training_step(gb, 0))
    assert abs(l0 - want_loss) <= LOSS_TOL * abs(want_loss)
