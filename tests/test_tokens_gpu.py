"""Token assembly and the fused action heads + loss (csrc/tokens.cu) against plain torch
restatements of the reference code: `ACTPCD.coord_embedding_sine` + the flatten / permute / cat of
`Transformer.forward` (act.py:467-506, transformer.py:75-92) and `forward_decoder` / `forward_loss`
(act.py:255-291, RLBench :770-825)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _policy(hidden=128, rlbench=False, **kw):
    from pointcloudmatters_b200.act import build_policy

    cfg = dict(hidden_dim=hidden, nhead=hidden // 64, dim_feedforward=32, enc_layers=1, dec_layers=1, dropout=0.0, num_queries=12,
               action_dim=11 if rlbench else 7, qpos_dim=9, goal_cond_dim=3, latent_dim=32, kl_weight=10.0, pcd_npoints=64,
               pcd_nsample=16, collision=True, position_loss_weight=3.0)
    cfg.update(kw)
    torch.manual_seed(0)
    return build_policy(cfg, rlbench=rlbench).cuda().train(), cfg


@pytest.mark.parametrize("hidden,b,m", [(128, 3, 64), (512, 4, 512), (96, 2, 40)])
def test_sine_embedding_tokens_match_the_eager_embedding(hidden, b, m):
    policy, _ = _policy(hidden if hidden % 64 == 0 else 128)
    policy.hidden_dim = hidden
    policy.additional_pos_embed = torch.nn.Embedding(3, hidden).cuda()
    coord = (torch.rand(b * m, 3, device="cuda") - 0.5) * 2.0
    want = policy.coord_embedding_sine(coord)  # (b*m, hidden), the reference's formula in torch ops
    pos = policy.coord_embedding_sine_tokens(coord, b)
    head = policy._n_head_rows()
    assert pos.shape == (head + m, b, hidden)
    got = pos[head:].permute(1, 0, 2).reshape(b * m, hidden)
    torch.testing.assert_close(got, want, rtol=0, atol=2e-7)
    add = policy.additional_pos_embed.weight.detach()
    assert torch.equal(pos[:head], add[:, None, :].expand(head, b, hidden))
    npf = hidden // 3
    assert float(got[:, 3 * npf:].abs().max()) == 0.0 if hidden > 3 * npf else True


@pytest.mark.parametrize("rlbench", [False, True])
def test_fused_heads_and_loss_match_torch(rlbench):
    from pointcloudmatters_b200 import functional as PF

    policy, cfg = _policy(128, rlbench)
    B, Q, E, A = 5, 12, 128, cfg["action_dim"]
    g = torch.Generator().manual_seed(3)
    hs_mem = torch.randn(Q, B, E, generator=g).cuda().requires_grad_(True)  # the decoder's memory order
    actions = torch.randn(B, Q, A, generator=g).cuda()
    is_pad = (torch.rand(B, Q, generator=g) < 0.3).cuda()
    mu = torch.randn(B, 32, generator=g).cuda().requires_grad_(True)
    logvar = (0.3 * torch.randn(B, 32, generator=g)).cuda().requires_grad_(True)
    with torch.no_grad():
        policy.action_head.bias.normal_(0, 0.1)
        policy.is_pad_head.bias.normal_(0, 0.1)
    sig, n_pos, w_pos = policy._head_cfg()

    def torch_ref(hs):
        a = F.linear(hs, policy.action_head.weight, policy.action_head.bias)
        if sig < A:
            a = torch.cat([a[..., :sig], torch.sigmoid(a[..., sig:])], -1)
        pad_hat = F.linear(hs, policy.is_pad_head.weight, policy.is_pad_head.bias)
        l = F.mse_loss(a, actions, reduction="none")
        if n_pos:
            l = torch.cat([l[..., :n_pos] * w_pos, l[..., n_pos:]], -1)
        action_loss = (l * ~is_pad.unsqueeze(-1)).mean()
        kl = (-0.5 * (1 + logvar - mu.pow(2) - logvar.exp())).sum(1).mean(0, True)[0]
        return a, pad_hat, action_loss + kl * 10.0, action_loss, kl

    hs = hs_mem.transpose(0, 1)  # (B, Q, E) view with the decoder's strides
    want = torch_ref(hs)
    (want[2] + 0.5 * want[3]).backward()
    ref_grads = [t.grad.clone() for t in (hs_mem, mu, logvar, policy.action_head.weight, policy.action_head.bias)]
    for t in (hs_mem, mu, logvar, policy.action_head.weight, policy.action_head.bias):
        t.grad = None
    assert policy.is_pad_head.weight.grad is None
    got = PF.act_heads_loss(hs, policy.action_head, policy.is_pad_head, actions, is_pad, mu, logvar, 10.0, sig, n_pos, w_pos)
    for a, b_ in zip(got, want):
        torch.testing.assert_close(a, b_.detach(), rtol=2e-5, atol=2e-6)
    (got[2] + 0.5 * got[3]).backward()
    grads = [t.grad for t in (hs_mem, mu, logvar, policy.action_head.weight, policy.action_head.bias)]
    for a, b_ in zip(grads, ref_grads):
        torch.testing.assert_close(a, b_, rtol=1e-4, atol=1e-7)
    assert policy.is_pad_head.weight.grad is None and policy.is_pad_head.bias.grad is None  # never receives one
    # a second call right away: the self-cleaning workspace must be zero again
    again = PF.act_heads_loss(hs, policy.action_head, policy.is_pad_head, actions, is_pad, mu, logvar, 10.0, sig, n_pos, w_pos)
    torch.testing.assert_close(again[2], got[2].detach(), rtol=1e-6, atol=0)
    # heads only (inference) + gradients arriving at the head outputs themselves
    a_hat, pad_hat, l0, _, _ = PF.act_heads_loss(hs, policy.action_head, policy.is_pad_head, None, None, None, None, 10.0, sig, n_pos, w_pos)
    assert l0 is None
    torch.testing.assert_close(a_hat, want[0].detach(), rtol=2e-5, atol=2e-6)
    for t in (hs_mem, policy.action_head.weight):
        t.grad = None
    ga, gp = torch.randn_like(a_hat), torch.randn_like(pad_hat)
    torch.autograd.backward([a_hat, pad_hat], [ga, gp])
    got_h, got_w, got_p = hs_mem.grad.clone(), policy.action_head.weight.grad.clone(), policy.is_pad_head.weight.grad.clone()
    for t in (hs_mem, policy.action_head.weight, policy.is_pad_head.weight, policy.is_pad_head.bias, policy.action_head.bias):
        t.grad = None
    w = torch_ref(hs_mem.transpose(0, 1))
    torch.autograd.backward([w[0], w[1]], [ga, gp])
    torch.testing.assert_close(got_h, hs_mem.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(got_w, policy.action_head.weight.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got_p, policy.is_pad_head.weight.grad, rtol=1e-4, atol=1e-5)


def test_token_fast_path_equals_the_concatenating_path():
    """The same policy run twice on the same batch: token buffers written directly by the set-abstraction head / sine
    kernel vs. the reference-shaped (b, c, 1, n) tensors concatenated inside Transformer.forward.  Same kernels on the
    same values downstream, so outputs and gradients agree to reassociation noise."""
    from pointcloudmatters_b200._lib import lib
    from pointcloudmatters_b200.data import synthetic_act_batch, to_device

    policy, cfg = _policy(128, enc_layers=2, dec_layers=2)
    h = synthetic_act_batch(4, 256, num_queries=12, seed=5)
    results = []
    for fast in (True, False):
        policy.zero_grad(set_to_none=True)
        policy.transformer.accepts_token_buffers = fast
        b = to_device(h, "cuda")
        b["pcds"]["n_max"] = h["pcds"]["n_max"]
        b["_eps"] = torch.randn(4, 32, generator=torch.Generator().manual_seed(1)).cuda()
        n0 = lib.calls.get("pcm_fill_head_rows", 0)
        out = policy(b)
        out["loss"].backward()
        assert (lib.calls.get("pcm_fill_head_rows", 0) > n0) == fast
        results.append((out["loss"].detach().clone(), out["a_hat"].detach().clone(),
                        {k: p.grad.clone() for k, p in policy.named_parameters() if p.grad is not None}))
    (l1, a1, g1), (l2, a2, g2) = results
    torch.testing.assert_close(l1, l2, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(a1, a2, rtol=1e-3, atol=1e-4)
    assert g1.keys() == g2.keys()
    scale = max(float(v.norm()) for v in g2.values())
    for k in g1:
        assert float((g1[k] - g2[k]).norm()) <= 2e-3 * max(float(g2[k].norm()), 1e-3 * scale), k


def test_grouped_memory_kv_equals_per_layer_projection():
    """Decoder cross-attention keys / values projected for all layers in ONE launch (functional.memory_kv) vs. inside each
    layer (the reference's structure, transformer.py:317-346): same outputs and the same gradients for every parameter,
    the memory and the learned positional rows."""
    from pointcloudmatters_b200._lib import lib
    from pointcloudmatters_b200.data import synthetic_act_batch, to_device

    policy, cfg = _policy(128, enc_layers=1, dec_layers=3)
    h = synthetic_act_batch(4, 256, num_queries=12, seed=8)
    results = []
    for grouped in (True, False):
        policy.zero_grad(set_to_none=True)
        policy.transformer.decoder.group_memory_kv = grouped
        b = to_device(h, "cuda")
        b["pcds"]["n_max"] = h["pcds"]["n_max"]
        b["_eps"] = torch.randn(4, 32, generator=torch.Generator().manual_seed(2)).cuda()
        n0 = lib.calls.get("pcm_gather_slices", 0)
        out = policy(b)
        # make every decoder layer live for this comparison: add the later intermediate outputs to the loss
        hs_all = policy.transformer.decoder  # noqa: F841  (layers 1.. are dead for the ACT loss itself)
        out["loss"].backward()
        assert (lib.calls.get("pcm_gather_slices", 0) > n0) == grouped
        results.append((out["loss"].detach().clone(), out["a_hat"].detach().clone(),
                        {k: p.grad.clone() for k, p in policy.named_parameters() if p.grad is not None}))
    (l1, a1, g1), (l2, a2, g2) = results
    torch.testing.assert_close(l1, l2, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(a1, a2, rtol=1e-3, atol=1e-4)
    assert g1.keys() == g2.keys()
    scale = max(float(v.norm()) for v in g2.values())
    for k in g1:
        assert float((g1[k] - g2[k]).norm()) <= 2e-3 * max(float(g2[k].norm()), 1e-3 * scale), k


def test_grouped_memory_kv_with_all_layers_live():
    """Same comparison on the bare Transformer with a loss over ALL intermediate decoder outputs, so that every layer's
    dK / dV block of the shared buffer carries a real gradient."""
    from pointcloudmatters_b200.transformer import Transformer

    torch.manual_seed(4)
    tr = Transformer(d_model=128, nhead=2, num_encoder_layers=1, num_decoder_layers=3, dim_feedforward=32, dropout=0.0,
                     return_intermediate_dec=True).cuda().train()
    B, M, E, Q = 3, 70, 128, 9
    g = torch.Generator().manual_seed(0)
    src = torch.randn(B, E, 1, M, generator=g).cuda().requires_grad_(True)
    pos = torch.randn(B, E, 1, M, generator=g).cuda()
    qe = torch.randn(Q, E, generator=g).cuda().requires_grad_(True)
    lat = torch.randn(1, B, E, generator=g).cuda().requires_grad_(True)
    prop = torch.randn(2, B, E, generator=g).cuda()
    add = torch.randn(3, E, generator=g).cuda().requires_grad_(True)
    wts = torch.randn(3, B, Q, E, generator=g).cuda()
    res = []
    for grouped in (True, False):
        for t in (src, qe, lat, add):
            t.grad = None
        tr.zero_grad(set_to_none=True)
        tr.decoder.group_memory_kv = grouped
        hs = tr(src, None, qe, pos, lat, prop, add)
        (hs * wts).sum().backward()
        res.append((hs.detach().clone(), [t.grad.clone() for t in (src, qe, lat, add)],
                    {k: p.grad.clone() for k, p in tr.named_parameters()}))
    (h1, i1, p1), (h2, i2, p2) = res
    torch.testing.assert_close(h1, h2, rtol=1e-3, atol=1e-4)
    for a, b_ in zip(i1, i2):
        assert float((a - b_).norm()) <= 3e-3 * float(b_.norm()) + 1e-6
    scale = max(float(v.norm()) for v in p2.values())
    for k in p1:
        assert float((p1[k] - p2[k]).norm()) <= 3e-3 * max(float(p2[k].norm()), 1e-3 * scale), k
