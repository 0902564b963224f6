"""Fused tcgen05 attention kernels (csrc/flash_attn.cu, through the C ABI) against an fp32 torch
restatement of nn.MultiheadAttention's math path (transformer.py:246-248,329-340: scaled scores,
key-padding mask, softmax, dropout, value product) on the same bf16-representable inputs.
Tolerances (bf16 operands, bf16 probabilities, fp32 accumulation): O rel-L2 <= 1e-2, gradients
rel-L2 <= 2e-2.  Dropout is checked EXACTLY: the kernel's counter-based keep-mask is replicated on
the host (tests/_dropout_mask.py) and fed to the torch reference."""
import numpy as np
import pytest
import torch

from tests._dropout_mask import keep_mask, keep_scale

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


def _run(B, nh, L, S, mask, p, seed=0):
    from pointcloudmatters_b200 import kernels as K

    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(seed)
    Z, E = B * nh, nh * 64
    q = torch.randn(Z, L, 64, device=dev, generator=g).bfloat16()
    k = torch.randn(Z, S, 64, device=dev, generator=g).bfloat16()
    v = torch.randn(Z, S, 64, device=dev, generator=g).bfloat16()
    do = torch.randn(Z, L, 64, device=dev, generator=g).bfloat16()
    kpm = None
    if mask:
        kpm = torch.zeros(B, S, dtype=torch.uint8, device=dev)
        kpm[:, max(S - 5, 1):] = 1
        kpm[0, S // 2] = 1
    seed_base = torch.tensor([123456789 + seed], dtype=torch.int64, device=dev) if p > 0 else None
    seed_off = 0x100000001B3 * 3
    O, lse = K.flash_attn_fwd(q.view(Z * L, 64), k.view(Z * S, 64), v.view(Z * S, 64), B, nh, L, S, kpm, 0.125, p,
                              seed_base, seed_off)
    buf = torch.full((L * B, E), float("nan"), dtype=torch.bfloat16, device=dev)
    kvbuf = torch.full((S * B, 2 * E), float("nan"), dtype=torch.bfloat16, device=dev)
    K.flash_attn_bwd(q.view(Z * L, 64), k.view(Z * S, 64), v.view(Z * S, 64), O, do.view(Z * L, 64), lse, B, nh, L, S, kpm,
                     0.125, p, seed_base, seed_off, buf, kvbuf[:, :E], kvbuf[:, E:])
    torch.cuda.synchronize()
    # fp32 reference
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    s = qf @ kf.transpose(1, 2) * 0.125
    if kpm is not None:
        s = s.view(B, nh, L, S).masked_fill(kpm.bool().view(B, 1, 1, S), float("-inf")).view(Z, L, S)
    a = torch.softmax(s, -1)
    if p > 0:
        keep = torch.from_numpy(keep_mask(int(seed_base.item()), seed_off, Z, L, S, p)).to(dev)
        frac = 1.0 - keep.float().mean().item()
        assert abs(frac - p) < 0.02, frac
        a = a * keep * keep_scale(p)
    o = a @ vf
    o.backward(do.float())
    tok = lambda t, n: t.view(B, nh, n, 64).permute(2, 0, 1, 3).reshape(n * B, E)  # head-merge
    ref_lse = torch.logsumexp(s, -1) * 1.4426950408889634
    return (O.float(), tok(o.detach(), L)), (lse, ref_lse.detach()), (buf.float(), tok(qf.grad, L)), \
        (kvbuf[:, :E].float(), tok(kf.grad, S)), (kvbuf[:, E:].float(), tok(vf.grad, S))


@pytest.mark.parametrize("B,nh,L,S,mask,p", [
    (2, 2, 128, 128, False, 0.0),      # exactly one tile each way
    (2, 8, 515, 515, False, 0.0),      # cfg-2 encoder self-attention (tail of 3)
    (2, 8, 100, 515, False, 0.0),      # cfg-2 decoder cross-attention
    (3, 2, 102, 102, True, 0.0),       # CVAE encoder with key-padding mask
    (2, 2, 100, 100, False, 0.0),      # decoder self-attention
    (1, 2, 7, 130, True, 0.0),         # tiny ragged
    (1, 8, 2051, 2051, False, 0.0),    # cfg-4 encoder (17 tiles)
    (2, 2, 300, 259, True, 0.1),       # dropout + mask, ragged both ways
    (2, 8, 515, 515, False, 0.1),      # cfg-2 with dropout
    (1, 1, 1, 1, False, 0.0),
])
def test_flash_attention_matches_reference(B, nh, L, S, mask, p):
    (O, rO), (lse, rlse), (dq, rdq), (dk, rdk), (dv, rdv) = _run(B, nh, L, S, mask, p)
    assert torch.isfinite(O).all() and torch.isfinite(dq).all() and torch.isfinite(dk).all() and torch.isfinite(dv).all()
    assert _rel(O, rO) <= 1e-2, _rel(O, rO)
    torch.testing.assert_close(lse, rlse, rtol=1e-3, atol=2e-2)
    for name, got, ref in (("dQ", dq, rdq), ("dK", dk, rdk), ("dV", dv, rdv)):
        assert _rel(got, ref) <= 2e-2, (name, _rel(got, ref))


def test_flash_attention_fully_masked_row_is_zero():
    """All keys padded for one batch element: the reference softmax yields NaN there; the kernels
    return zeros (documented divergence, never reached on the ACT path)."""
    from pointcloudmatters_b200 import kernels as K

    B, nh, L, S = 2, 1, 40, 70
    q = torch.randn(B * nh * L, 64, device="cuda").bfloat16()
    k = torch.randn(B * nh * S, 64, device="cuda").bfloat16()
    v = torch.randn(B * nh * S, 64, device="cuda").bfloat16()
    kpm = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    kpm[1] = 1
    O, lse = K.flash_attn_fwd(q, k, v, B, nh, L, S, kpm, 0.125, 0.0, None, 0)
    O = O.view(L, B, 64)
    assert torch.isfinite(O[:, 0]).all() and (O[:, 1] == 0).all()


def test_flash_attention_growing_scores_take_exact_rescale_path():
    """Later key tiles carry scores thousands of octaves above the first tile's maximum: the lazy
    reference maximum of the forward kernel overflows there and the warp must fall back to the
    exact max + rescale path (and the backward must stay finite with the saved log-sum-exp)."""
    from pointcloudmatters_b200 import kernels as K

    B, nh, L, S = 1, 2, 200, 400
    Z, E = B * nh, nh * 64
    g = torch.Generator(device="cuda").manual_seed(11)
    q = (4 * torch.randn(Z, L, 64, device="cuda", generator=g)).bfloat16()
    k = torch.randn(Z, S, 64, device="cuda", generator=g)
    k[:, 128:] *= 200.0
    k[:, 300:] *= 0.01  # and back down again: the reference must not be lowered
    k = k.bfloat16()
    v = torch.randn(Z, S, 64, device="cuda", generator=g).bfloat16()
    O, lse = K.flash_attn_fwd(q.view(Z * L, 64), k.view(Z * S, 64), v.view(Z * S, 64), B, nh, L, S, None, 0.125, 0.0, None, 0)
    s = q.float() @ k.float().transpose(1, 2) * 0.125
    ref = torch.softmax(s, -1) @ v.float()
    ref_tok = ref.view(B, nh, L, 64).permute(2, 0, 1, 3).reshape(L * B, E)
    assert torch.isfinite(O.float()).all()
    assert _rel(O.float(), ref_tok) <= 1e-2, _rel(O.float(), ref_tok)
    torch.testing.assert_close(lse, torch.logsumexp(s, -1) * 1.4426950408889634, rtol=1e-3, atol=5e-2)
    do = torch.randn(Z * L, 64, device="cuda", generator=g).bfloat16()
    dq = torch.empty(L * B, E, dtype=torch.bfloat16, device="cuda")
    dkv = torch.empty(S * B, 2 * E, dtype=torch.bfloat16, device="cuda")
    K.flash_attn_bwd(q.view(Z * L, 64), k.view(Z * S, 64), v.view(Z * S, 64), O, do, lse, B, nh, L, S, None, 0.125, 0.0, None, 0,
                     dq, dkv[:, :E], dkv[:, E:])
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
