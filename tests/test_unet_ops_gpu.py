"""Operator-level parity of the Diffusion-Policy denoiser kernels (csrc/unet1d.cu through the C ABI,
pointcloudmatters_b200/functional_unet.py) against plain fp32 torch on the same inputs.

Tolerances: GroupNorm+Mish / Mish / unfold / fold are fp32 kernels -> 1e-5 relative; convolutions run their
contraction on bf16 tensor-core operands with fp32 accumulation -> rel-L2 <= 1e-2 against fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


@pytest.mark.parametrize("B,T,C,G,film,res", [(3, 16, 32, 8, True, False), (2, 8, 64, 8, False, True),
                                              (5, 4, 2048, 8, True, True), (2, 16, 24, 3, False, False),
                                              (1, 16, 512, 8, True, True)])
def test_groupnorm_mish_matches_torch(B, T, C, G, film, res):
    from pointcloudmatters_b200 import functional_unet as UF

    g = torch.Generator(device="cuda").manual_seed(B * 100 + C)
    x = torch.randn(B, T, C, device="cuda", generator=g) * 2 + 0.5
    gn = torch.nn.GroupNorm(G, C).cuda()
    with torch.no_grad():
        gn.weight.add_(0.3 * torch.randn(C, device="cuda", generator=g))
        gn.bias.add_(0.3 * torch.randn(C, device="cuda", generator=g))
    fl = torch.randn(B, 2 * C, device="cuda", generator=g) if film else None
    rs = torch.randn(B, T, C, device="cuda", generator=g) if res else None
    dy = torch.randn(B, T, C, device="cuda", generator=g)

    def run(fn):
        leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (x, fl, rs)]
        gn.zero_grad(set_to_none=True)
        y = fn(*leaves)
        y.backward(dy)
        return y.detach(), [t.grad if t is not None else None for t in leaves], gn.weight.grad.clone(), gn.bias.grad.clone()

    def ref(x_, fl_, rs_):
        y = F.mish(F.group_norm(x_.permute(0, 2, 1), G, gn.weight, gn.bias, gn.eps)).permute(0, 2, 1)
        if fl_ is not None:
            y = fl_[:, None, :C] * y + fl_[:, None, C:]
        return y + rs_ if rs_ is not None else y

    y0, g0, dg0, db0 = run(ref)
    y1, g1, dg1, db1 = run(lambda a, b, c: UF.groupnorm_mish(a, gn, b, c))
    assert _rel(y1, y0) < 1e-5
    assert _rel(y1._pcm_bf16.float() if hasattr(y1, "_pcm_bf16") else y1, y0) < 1e-2
    for a, b in zip(g1, g0):
        if b is not None:
            assert _rel(a, b) < 2e-5
    assert _rel(dg1, dg0) < 2e-5 and _rel(db1, db0) < 2e-5


@pytest.mark.parametrize("B,L,Cin,Cout,k,stride,pad", [(3, 16, 7, 32, 5, 1, 2), (2, 16, 32, 64, 5, 1, 2),
                                                       (4, 16, 64, 64, 3, 2, 1), (2, 8, 128, 7, 1, 1, 0),
                                                       (2, 4, 256, 128, 1, 1, 0), (2, 5, 16, 24, 5, 1, 2)])
def test_conv1d_channel_last_matches_torch(B, L, Cin, Cout, k, stride, pad):
    from pointcloudmatters_b200 import functional_unet as UF

    g = torch.Generator(device="cuda").manual_seed(L * 10 + Cin)
    conv = torch.nn.Conv1d(Cin, Cout, k, stride, pad).cuda()
    x = torch.randn(B, L, Cin, device="cuda", generator=g)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y0 = conv(xa.permute(0, 2, 1)).permute(0, 2, 1)
    dy = torch.randn(y0.shape, device="cuda", generator=g)
    y0.backward(dy)
    gw0, gb0 = conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.zero_grad(set_to_none=True)
    y1 = UF.conv1d_cl(xb, conv.weight, conv.bias, stride, pad)
    assert y1.shape == y0.shape
    y1.backward(dy)
    assert _rel(y1, y0) < 1e-2
    assert _rel(xb.grad, xa.grad) < 1e-2
    assert _rel(conv.weight.grad, gw0) < 1e-2 and _rel(conv.bias.grad, gb0) < 1e-2


@pytest.mark.parametrize("B,L,C", [(3, 4, 64), (2, 8, 32), (1, 5, 16)])
def test_conv_transpose1d_channel_last_matches_torch(B, L, C):
    from pointcloudmatters_b200 import functional_unet as UF

    g = torch.Generator(device="cuda").manual_seed(L + C)
    conv = torch.nn.ConvTranspose1d(C, C, 4, 2, 1).cuda()
    x = torch.randn(B, L, C, device="cuda", generator=g)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y0 = conv(xa.permute(0, 2, 1)).permute(0, 2, 1)
    dy = torch.randn(y0.shape, device="cuda", generator=g)
    y0.backward(dy)
    gw0, gb0 = conv.weight.grad.clone(), conv.bias.grad.clone()
    conv.zero_grad(set_to_none=True)
    y1 = UF.conv_transpose1d_cl(xb, conv.weight, conv.bias, 2, 1)
    assert y1.shape == y0.shape == (B, 2 * L, C)
    y1.backward(dy)
    assert _rel(y1, y0) < 1e-2
    assert _rel(xb.grad, xa.grad) < 1e-2
    assert _rel(conv.weight.grad, gw0) < 1e-2 and _rel(conv.bias.grad, gb0) < 1e-2


def test_unfold_fold_are_exact_adjoints():
    """unfold is a 0/1 gather and fold its transpose: <unfold(x), c> == <x, fold(c)> up to bf16 rounding of
    unfold's output, and fold(unfold(1)) counts the taps covering each position."""
    from pointcloudmatters_b200 import functional_unet as UF

    B, L, C, k, s, p = 2, 16, 8, 5, 1, 2
    R = (L + 2 * p - k) // s + 1
    x = torch.ones(B, L, C, device="cuda")
    col = UF._unfold(x, k, s, p, R).float()
    cover = UF._fold(col.contiguous(), B, L, C, k, s, p, R)
    t = torch.arange(L, device="cuda")
    want = (torch.minimum(t, torch.tensor(2, device="cuda")) + torch.minimum(L - 1 - t, torch.tensor(2, device="cuda")) + 1).float()
    assert torch.equal(cover, want[None, :, None].expand(B, L, C))


def test_mish_matches_torch():
    from pointcloudmatters_b200 import functional_unet as UF

    x = torch.linspace(-30, 30, 4001, device="cuda").requires_grad_(True)
    y = UF.mish(x)
    y.sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    y2 = F.mish(x2)
    y2.sum().backward()
    assert torch.allclose(y, y2, rtol=1e-5, atol=1e-6)
    assert torch.allclose(x.grad, x2.grad, rtol=1e-4, atol=1e-6)


def test_empty_and_degenerate_shapes_through_the_abi():
    """Edge cases at the C-ABI level: empty batches are no-ops (status 0), single time step / single group work,
    invalid geometry is rejected with a negative status instead of launching."""
    from pointcloudmatters_b200 import functional_unet as UF
    from pointcloudmatters_b200._lib import current_stream, lib, ptr

    st = current_stream()
    x = torch.zeros(4, device="cuda")
    assert lib.pcm_conv1d_unfold(0, 16, 8, 5, 1, 2, 16, ptr(x), 0, 8, ptr(x), 40, st) == 0
    assert lib.pcm_conv1d_fold(0, 16, 8, 5, 1, 2, 16, ptr(x), 40, None, ptr(x), None, st) == 0
    assert lib.pcm_groupnorm_mish_fwd(0, 16, 8, 2, ptr(x), ptr(x), ptr(x), 1e-5, None, None, ptr(x), None, ptr(x), ptr(x), st) == 0
    assert lib.pcm_mish_fwd(0, ptr(x), ptr(x), None, st) == 0
    assert lib.pcm_bn_stats(0, 8, ptr(x), ptr(x), st) == 0
    assert lib.pcm_conv1d_unfold(1, 16, 8, 5, 1, 2, 16, ptr(x), 0, 8, ptr(x), 8, st) < 0      # ldc < C*k
    assert lib.pcm_groupnorm_mish_fwd(1, 16, 10, 4, ptr(x), ptr(x), ptr(x), 1e-5, None, None, ptr(x), None, ptr(x), ptr(x), st) < 0
    assert lib.pcm_bn_stats(4, 6, ptr(x), ptr(x), st) < 0                                     # C % 4
    # T = 1 (a single time step), one group, odd length with stride 2
    gn = torch.nn.GroupNorm(1, 8).cuda()
    xs = torch.randn(3, 1, 8, device="cuda")
    want = F.mish(F.group_norm(xs.permute(0, 2, 1), 1, gn.weight, gn.bias, gn.eps)).permute(0, 2, 1)
    assert torch.allclose(UF.groupnorm_mish(xs, gn), want, rtol=1e-5, atol=1e-6)
    conv = torch.nn.Conv1d(8, 16, 3, 2, 1).cuda()
    xo = torch.randn(2, 7, 8, device="cuda")
    got = UF.conv1d_cl(xo, conv.weight, conv.bias, 2, 1)
    ref = conv(xo.permute(0, 2, 1)).permute(0, 2, 1)
    assert got.shape == ref.shape == (2, 4, 16)
    assert float((got - ref).norm() / ref.norm()) < 1e-2
