"""BCTrainer behaviour that needs the CUDA step: gradient accumulation, optimizer checkpoint round trip,
load_state_dict resynchronising the bf16 operand copy, the bounded CUDA-graph cache with its eager fallback, and the
debug validation of the cloud-size hints."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(hidden_dim=128, nhead=2, dim_feedforward=32, enc_layers=1, dec_layers=2, dropout=0.0, num_queries=12,
           action_dim=7, qpos_dim=9, goal_cond_dim=3, latent_dim=32, kl_weight=10.0, pcd_npoints=64, pcd_nsample=16)


def _batch(seed, b=4, n=256, ragged=False):
    from pointcloudmatters_b200.data import synthetic_act_batch, to_device

    h = synthetic_act_batch(b, n, num_queries=12, seed=seed, ragged=ragged)
    g = to_device(h, "cuda")
    g["pcds"]["n_max"] = h["pcds"]["n_max"]
    g["_eps"] = torch.randn(b, 32, generator=torch.Generator().manual_seed(seed)).cuda()
    return g


def _module(seed=0, **kw):
    from pointcloudmatters_b200.act import build_policy
    from pointcloudmatters_b200.bc_module import ACTBCModule

    torch.manual_seed(seed)
    m = ACTBCModule(build_policy(CFG).cuda().train(), optimizer=dict(lr=1e-3), total_steps=50, **kw)
    m.configure_optimizers()
    return m


def test_gradient_accumulation_sums_micro_batches_and_steps_once():
    """accumulate_grad_batches=2 (reference preset, exp_maniskill2_act_policy/base.yaml:26-28): the flat gradient holds
    g(b1) + g(b2) when the group closes, ONE optimizer step is taken with grad_scale 1/2."""
    b0, b1, b2 = _batch(1), _batch(2), _batch(3)
    single = {}
    for name, b in (("b1", b1), ("b2", b2)):
        m = _module()
        m.training_step(b0, 0)  # builds the flat state (same weights afterwards in every replica: same seed, same batch)
        tr = m._trainer
        tr._forward_backward(b)
        single[name] = tr.flat.grad[: tr.flat.n_active].clone()
    m = _module(accumulate_grad_batches=2)
    tr = m._trainer
    assert tr.hyper_values(0)[8] == 0.5
    m.training_step(b0, 0)
    m.training_step(b0, 1)  # closes the first group -> weights equal the replicas above?  no: two micro-batches were used
    assert tr.step_num == 1
    # fresh module again, this time compare the accumulated buffer itself
    m = _module(accumulate_grad_batches=2)
    tr = m._trainer
    snap = {}
    orig = tr.reduce_gradients
    tr.reduce_gradients = lambda: (snap.setdefault("g", tr.flat.grad[: tr.flat.n_active].clone()), orig())[1]
    m2 = _module()
    m2.training_step(b0, 0)
    # bring `m` to the same weights as the single-step replicas: one closed group of (b0, b0) differs from one step on
    # b0, so copy the weights over instead
    m.training_step(b0, 0)
    m.training_step(b0, 1)
    snap.clear()
    with torch.no_grad():
        tr.flat.param.copy_(m2._trainer.flat.param)
        tr.flat.sync_shadow()
    steps_before = tr.step_num
    m.training_step(b1, 2)
    assert tr.step_num == steps_before and "g" not in snap  # group still open: no all-reduce, no optimizer step
    m.training_step(b2, 3)
    assert tr.step_num == steps_before + 1
    want = single["b1"] + single["b2"]
    err = float((snap["g"] - want).norm() / want.norm())
    assert err <= 2e-3, err  # fp32 atomics reorder sums between runs


def test_optimizer_state_round_trip_and_shadow_resync():
    from pointcloudmatters_b200.act import build_policy
    from pointcloudmatters_b200.bc_module import ACTBCModule

    a = _module()
    for i in range(3):
        a.training_step(_batch(10 + i), i)
    sd_w = {k: v.detach().clone() for k, v in a.policy.state_dict().items()}
    sd_o = a._trainer.state_dict()
    assert sd_o["step_num"] == 3 and "is_pad_head.weight" in sd_o["inactive"]
    assert set(sd_o["exp_avg"]) == {n for n, _ in a.policy.named_parameters()} - set(sd_o["inactive"])
    torch.manual_seed(123)  # different init on purpose
    b = ACTBCModule(build_policy(CFG).cuda().train(), optimizer=dict(lr=1e-3), total_steps=50)
    b.configure_optimizers()
    b.policy.load_state_dict(sd_w)
    b._trainer.load_state_dict(sd_o)
    assert b._trainer.step_num == 3 and b._trainer.flat is not None
    nxt = _batch(20)
    la, lb = float(a.training_step(nxt, 3)), float(b.training_step(nxt, 3))
    assert abs(la - lb) <= 2e-3 * abs(la)
    pa, pb = a._trainer.flat, b._trainer.flat
    na = {n: p for n, p in a.policy.named_parameters()}
    for n, p in b.policy.named_parameters():
        assert float((p - na[n]).abs().max()) <= 2e-3 * 1e-3 + 1e-6 + 2e-3 * float(na[n].abs().max()) * 1e-2, n
    # policy.load_state_dict AFTER the flat state exists must refresh the bf16 operand copy (post hook)
    with torch.no_grad():
        changed = {k: (v * 1.5 if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd_w.items()}
    b.policy.load_state_dict(changed)
    f = b._trainer.flat
    assert torch.equal(f.param_bf16, f.param.to(torch.bfloat16))


def test_graph_cache_is_bounded_and_falls_back_when_shapes_never_repeat():
    m = _module(use_cuda_graph=True)
    tr = m._trainer
    tr.max_cached_graphs = 2
    for i in range(3):
        m.training_step(_batch(30 + i), i)  # eager warm-up steps, then the first capture
    for i, b in enumerate((4, 6, 8, 4)):
        m.training_step(_batch(40 + i, b=b), 3 + i)
        assert len(tr._graphs) <= 2
    assert tr.graph_disabled_reason is None
    # hints that differ by a few points share a graph (bucketed to 128)
    g1, g2 = _batch(50), _batch(51)
    g1["pcds"]["n_max"], g2["pcds"]["n_max"] = 250, 256
    assert tr._signature(tr._bucket_hints(tr._inputs_only(g1))) == tr._signature(tr._bucket_hints(tr._inputs_only(g2)))
    # ragged clouds: a new sum-N every batch -> the trainer gives up capturing instead of re-capturing forever
    last = None
    for i in range(24):
        last = float(m.training_step(_batch(60 + i, ragged=True), 10 + i))
    assert tr.graph_disabled_reason is not None and not tr._graphs and last == last


def test_debug_hints_reject_an_undersized_n_max():
    m = _module(debug_hints=True)
    b = _batch(70)
    b["pcds"]["n_max"] = 100  # clouds have 256 points
    with pytest.raises(ValueError):
        m.training_step(b, 0)


@pytest.mark.parametrize("use_graph", [False, True])
def test_training_step_accepts_a_pinned_host_batch(use_graph):
    """The batch may stay in pinned host memory: the step stages it host->device itself (graph path: straight into the
    captured graph's static inputs).  Same losses and weights as with a device-resident copy of the batch."""
    from pointcloudmatters_b200.data import synthetic_act_batch, to_device

    hosts = [synthetic_act_batch(4, 256, num_queries=12, seed=s, pin=True) for s in (5, 6)]
    for h in hosts:
        h["_eps"] = torch.randn(4, 32, generator=torch.Generator().manual_seed(7)).pin_memory()
    losses = {}
    for mode in ("device", "host"):
        m = _module(use_cuda_graph=use_graph)
        out = []
        for i in range(6):
            h = hosts[i % 2]
            if mode == "device":
                b = to_device(h, "cuda")
                b["pcds"]["n_max"] = h["pcds"]["n_max"]
            else:
                b = h
            out.append(float(m.training_step(b, i)))
        losses[mode] = out
        if use_graph:
            assert len(m._trainer._graphs) >= 1 and m._trainer.graph_disabled_reason is None
    assert losses["host"] == pytest.approx(losses["device"], rel=1e-2), losses  # (atomics noise, see the prefetch test)


def test_prefetch_stages_the_next_host_batch_and_changes_nothing():
    """BCTrainer.prefetch: the next step's pinned host batch is copied on a side stream while the current step runs; the
    step that consumes it gives the same losses as without prefetching (also when a prefetched batch is NOT the one used)."""
    from pointcloudmatters_b200.data import synthetic_act_batch

    hosts = [synthetic_act_batch(4, 256, num_queries=12, seed=s, pin=True) for s in (5, 6, 7)]
    for h in hosts:
        h["_eps"] = torch.randn(4, 32, generator=torch.Generator().manual_seed(7)).pin_memory()
    losses = {}
    for mode in ("plain", "prefetch"):
        m = _module(use_cuda_graph=True)
        out = []
        for i in range(9):
            loss = m.training_step(hosts[i % 3], i)
            if mode == "prefetch":
                nxt = hosts[(i + 1) % 3] if i != 5 else hosts[0]  # step 6 is handed a different batch than was prefetched
                m.prefetch(nxt)
            out.append(float(loss))
        losses[mode] = out
        assert len(m._trainer._graphs) >= 1 and m._trainer.graph_disabled_reason is None
        if mode == "prefetch":
            assert hasattr(m._trainer, "_stages") and len(m._trainer._stages) == 1
    # run-to-run noise of the fp32-atomic reductions reaches ~2e-3 after a few Adam steps (see test_act_gpu); a wrong or
    # stale batch would change the loss by tens of percent
    assert losses["prefetch"] == pytest.approx(losses["plain"], rel=1e-2), losses
