"""Multi-head attention operator (tcgen05 projection GEMMs + fused tcgen05 attention kernels) against
torch.nn.MultiheadAttention in fp32 on the same (bf16-representable) weights and inputs:
self- and cross-attention, key-padding mask, ragged (non-multiple-of-64) lengths, forward and all
gradients.  Tolerances (bf16 operands / bf16 probabilities): outputs rel-L2 <= 1.5e-2,
gradients rel-L2 <= 4e-2.  Dropout is checked exactly by replicating the kernel's counter-based keep-mask on the host."""
import pytest
import torch

from tests._dropout_mask import keep_mask, keep_scale

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


def _mk(E, nh, p=0.0, seed=0):
    torch.manual_seed(seed)
    m = torch.nn.MultiheadAttention(E, nh, dropout=p).cuda()
    with torch.no_grad():
        for q in m.parameters():
            q.copy_((q + 0.05 * torch.randn_like(q)).bfloat16().float())
    return m


@pytest.mark.parametrize("L,S,B,E,nh,mask,self_attn", [
    (102, 102, 3, 128, 2, True, True), (515, 515, 2, 512, 8, False, True), (100, 515, 2, 512, 8, False, False),
    (64, 64, 1, 128, 2, False, True), (7, 130, 2, 128, 2, True, False)])
def test_mha_matches_torch(L, S, B, E, nh, mask, self_attn):
    from pointcloudmatters_b200 import functional as PF

    mha = _mk(E, nh)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(L, B, E, device="cuda", generator=g).bfloat16().float()
    pos = torch.randn(L, B, E, device="cuda", generator=g).bfloat16().float() * 0.5
    mem = x if self_attn else torch.randn(S, B, E, device="cuda", generator=g).bfloat16().float()
    kpm = None
    if mask:
        kpm = torch.zeros(B, S, dtype=torch.bool, device="cuda")
        kpm[:, S - 5:] = True
        kpm[0, S // 2] = True
    dout = torch.randn(L, B, E, device="cuda", generator=g)
    res = []
    for ours in (False, True):
        xs = x.clone().requires_grad_(True)
        ms = xs if self_attn else mem.clone().requires_grad_(True)
        mha.zero_grad()
        if ours:
            out = (PF.multi_head_attention(mha, xs, pos, None, None, kpm, training=False) if self_attn
                   else PF.multi_head_attention(mha, xs, pos, ms, None, kpm, training=False))
        else:
            qk = xs + pos
            out = mha(qk, qk if self_attn else ms, xs if self_attn else ms, key_padding_mask=kpm)[0]
        out.backward(dout)
        res.append((out.detach(), xs.grad.clone(), None if self_attn else ms.grad.clone(),
                    mha.in_proj_weight.grad.clone(), mha.in_proj_bias.grad.clone(), mha.out_proj.weight.grad.clone(),
                    mha.out_proj.bias.grad.clone()))
    ref, got = res
    assert _rel(got[0], ref[0]) <= 1.5e-2
    names = ["dx", "dmem", "dW_in", "db_in", "dW_out", "db_out"]
    for i, nme in enumerate(names, start=1):
        if ref[i] is None:
            continue
        assert _rel(got[i], ref[i]) <= 4e-2, (nme, _rel(got[i], ref[i]))


def test_mha_dropout_is_consistent():
    """With p > 0: ~p of the probabilities are dropped, survivors are scaled by 1/(1-p), and the
    backward pass regenerates the same mask (gradients match a torch emulation using that mask)."""
    from pointcloudmatters_b200 import functional as PF

    E, nh, L, B, p = 128, 2, 96, 2, 0.25
    mha = _mk(E, nh, p=p)
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(L, B, E, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    out = PF.multi_head_attention(mha, x, None, None, None, None, training=True)
    fn = out.grad_fn
    seed_base, seed_off = fn.aux
    keep = torch.from_numpy(keep_mask(int(seed_base.item()), seed_off, B * nh, L, L, p)).cuda()
    frac = 1.0 - keep.float().mean().item()
    assert abs(frac - p) < 0.02, frac
    # torch emulation with the recovered mask
    dout = torch.randn(L, B, E, device="cuda", generator=g)
    out.backward(dout)
    gx = x.grad.clone()
    x2 = x.detach().clone().requires_grad_(True)
    W, bI = mha.in_proj_weight, mha.in_proj_bias
    q = (x2 @ W[:E].t() + bI[:E]).view(L, B, nh, 64).permute(1, 2, 0, 3)
    k = (x2 @ W[E:2 * E].t() + bI[E:2 * E]).view(L, B, nh, 64).permute(1, 2, 0, 3)
    v = (x2 @ W[2 * E:].t() + bI[2 * E:]).view(L, B, nh, 64).permute(1, 2, 0, 3)
    a = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) * keep.view(B, nh, L, L) * keep_scale(p)
    o = (a @ v).permute(2, 0, 1, 3).reshape(L, B, E) @ mha.out_proj.weight.t() + mha.out_proj.bias
    assert _rel(out.detach(), o.detach()) <= 1.5e-2
    o.backward(dout)
    assert _rel(gx, x2.grad) <= 4e-2


def test_weight_grads_accumulate_in_place_into_existing_buffers():
    """Second backward with populated .grad buffers takes the direct-accumulation path (kernels add
    into the parameters' own gradient memory, autograd receives None): gradients exactly double up
    to fp32 atomics ordering, and the buffers are not re-allocated.  Also covers the learned
    positional head (gradient only for the leading rows of a detached positional tensor)."""
    from pointcloudmatters_b200 import functional as PF

    E, nh, L, S, B = 128, 2, 70, 133, 3
    mha = _mk(E, nh)
    ln = torch.nn.LayerNorm(E).cuda()
    lin = torch.nn.Linear(E, 64).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(L, B, E, device="cuda", generator=g).bfloat16().float()
    mem = torch.randn(S, B, E, device="cuda", generator=g).bfloat16().float()
    head = torch.nn.Parameter(torch.randn(2, 1, E, device="cuda", generator=g).bfloat16().float())
    sine = torch.randn(S - 2, B, E, device="cuda", generator=g).bfloat16().float()
    params = list(mha.parameters()) + list(ln.parameters()) + list(lin.parameters()) + [head]

    def run(use_head):
        if use_head:
            mpos = torch.cat([head.detach().repeat(1, B, 1), sine], 0)
            a = PF.multi_head_attention(mha, x, None, mem, mpos, None, training=False, mem_pos_head=head)
        else:
            mpos = torch.cat([head.repeat(1, B, 1), sine], 0)
            a = PF.multi_head_attention(mha, x, None, mem, mpos, None, training=False)
        y = PF.add_dropout_layernorm(a, x, ln, 0.0, False)
        return PF.linear(y, lin.weight, lin.bias).square().mean()

    run(False).backward()
    ref = [p.grad.clone() for p in params]
    ptrs = [p.grad.data_ptr() for p in params]
    for p in params:
        p.grad = None
    run(True).backward()
    for p, r in zip(params, ref):
        assert _rel(p.grad, r) <= 1e-3, _rel(p.grad, r)
    ptrs = [p.grad.data_ptr() for p in params]
    run(True).backward()  # accumulates in place
    for p, r, q in zip(params, ref, ptrs):
        assert p.grad.data_ptr() == q
        assert _rel(p.grad, 2 * r) <= 1e-3, _rel(p.grad, 2 * r)
