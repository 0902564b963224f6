"""Fused residual + dropout + LayerNorm kernels and the column-sum kernel against plain PyTorch fp32
references of the same ops (fp32 arithmetic on both sides: rtol/atol 2e-5 forward, 1e-4 backward;
the column-wise gamma / beta / bias gradients are fp32 sums in a different order: 2e-4)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,C", [(1, 128), (1000, 512), (33000, 512), (77, 1024), (515, 256)])
@pytest.mark.parametrize("with_x", [True, False])
def test_add_dropout_ln_matches_torch(rows, C, with_x):
    from pointcloudmatters_b200 import functional as PF

    g = torch.Generator(device="cuda").manual_seed(rows + C)
    norm = torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        norm.weight.copy_(1 + 0.2 * torch.randn(C, device="cuda", generator=g))
        norm.bias.copy_(0.1 * torch.randn(C, device="cuda", generator=g))
    x0 = torch.randn(rows, C, device="cuda", generator=g) if with_x else None
    r0 = torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.3
    dy = torch.randn(rows, C, device="cuda", generator=g)
    outs = []
    for ours in (False, True):
        x = x0.clone().requires_grad_(True) if with_x else None
        r = r0.clone().requires_grad_(True)
        norm.zero_grad()
        if ours:
            y = PF.add_dropout_layernorm(x, r, norm, 0.1, training=False)
        else:
            y = F.layer_norm(r + x if with_x else r, (C,), norm.weight, norm.bias, norm.eps)
        y.backward(dy)
        outs.append((y.detach(), r.grad, x.grad if with_x else None, norm.weight.grad.clone(), norm.bias.grad.clone()))
    ref, got = outs
    torch.testing.assert_close(got[0], ref[0], rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(got[1], ref[1], rtol=1e-4, atol=1e-4)
    if with_x:
        torch.testing.assert_close(got[2], ref[2], rtol=1e-4, atol=1e-4)
    scale = max(1.0, (rows ** 0.5))
    torch.testing.assert_close(got[3], ref[3], rtol=2e-4, atol=2e-4 * scale)
    torch.testing.assert_close(got[4], ref[4], rtol=2e-4, atol=2e-4 * scale)


def test_add_dropout_ln_dropout_mask_is_consistent():
    from pointcloudmatters_b200 import functional as PF

    rows, C, p = 4000, 512, 0.1
    g = torch.Generator(device="cuda").manual_seed(0)
    norm = torch.nn.LayerNorm(C).cuda()
    x = (torch.randn(rows, C, device="cuda", generator=g) + 3.0).requires_grad_(True)  # no exact zeros
    r = torch.randn(rows, C, device="cuda", generator=g).requires_grad_(True)
    y = PF.add_dropout_layernorm(x, r, norm, p, training=True)
    h = y.grad_fn.saved_tensors[0]
    kept = (h - r.detach()) != 0
    frac = 1 - kept.float().mean().item()
    assert abs(frac - p) < 0.01, frac
    torch.testing.assert_close((h - r.detach())[kept], (x.detach() / (1 - p))[kept], rtol=1e-5, atol=1e-5)
    ref = F.layer_norm(r.detach() + x.detach() * kept / (1 - p), (C,), norm.weight, norm.bias, norm.eps)
    torch.testing.assert_close(y.detach(), ref, rtol=2e-5, atol=2e-5)
    dy = torch.randn(rows, C, device="cuda", generator=g)
    y.backward(dy)
    assert float(x.grad[~kept].abs().max()) == 0.0  # dropped elements get no gradient
    torch.testing.assert_close(x.grad[kept], (r.grad / (1 - p))[kept], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_ln_backward_bf16_only_dx_equals_the_fp32_form(p):
    """dx_fp32=False: the kernel skips the fp32 dx store; dres, bf16(dx), dgamma / dbeta and colsum(dx) are unchanged
    (up to the order of the atomics), and the placeholder is not touched."""
    from pointcloudmatters_b200 import kernels as K

    rows, C = 3000, 512
    g = torch.Generator(device="cuda").manual_seed(3)
    h = torch.randn(rows, C, device="cuda", generator=g)
    dy = torch.randn(rows, C, device="cuda", generator=g)
    gamma = torch.randn(C, device="cuda", generator=g)
    mean, rstd = h.mean(1), 1.0 / torch.sqrt(h.var(1, unbiased=False) + 1e-5)
    sb = torch.tensor([99], dtype=torch.int64, device="cuda")
    outs = []
    for fp32 in (True, False):
        cs = torch.zeros(C, device="cuda")
        dres, dx, dg, db, dxb = K.add_dropout_ln_bwd(dy, h, mean, rstd, gamma, p, sb if p > 0 else None, 5, True, want_dx_bf16=True,
                                                     dx_colsum=cs, dx_fp32=fp32)
        outs.append((dres, dx, dg, db, dxb, cs))
    a, b = outs
    assert torch.equal(a[0], b[0]) and torch.equal(a[4], b[4])
    assert torch.equal(a[4], a[1].bfloat16())
    assert b[1].data_ptr() != b[0].data_ptr()
    for i in (2, 3, 5):
        torch.testing.assert_close(a[i], b[i], rtol=1e-4, atol=1e-3)


def test_unwritten_dx_is_never_cast():
    """A consumer that misses the side-channel entry of a bf16-only dx must raise, not read the placeholder."""
    from pointcloudmatters_b200 import functional as PF
    from pointcloudmatters_b200._lib import PcmError

    C = 512
    norm = torch.nn.LayerNorm(C).cuda()
    bias = torch.nn.Parameter(torch.zeros(C, device="cuda"))
    x = torch.randn(4, 8, C, device="cuda", requires_grad=True)
    seen = {}

    class Probe(torch.autograd.Function):  # stands in for a producer that tags its output but loses the side channel
        @staticmethod
        def forward(ctx, t):
            return t.clone()

        @staticmethod
        def backward(ctx, gr):
            PF._GRAD_BF16.clear()
            try:
                PF._grad_bf16(gr, 32, C)
            except PcmError as e:
                seen["err"] = str(e)
            return torch.zeros_like(gr)

    xo = Probe.apply(x)
    xo._pcm_bias = bias
    y = PF.add_dropout_layernorm(xo, torch.randn(4, 8, C, device="cuda"), norm, 0.1, True, x_exclusive=True)
    y.sum().backward()
    assert "no fp32 data" in seen.get("err", "")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_colsum(dtype):
    from pointcloudmatters_b200.kernels import colsum

    g = torch.Generator(device="cuda").manual_seed(1)
    src = torch.randn(32960, 1024, device="cuda", generator=g).to(dtype)
    view = src[:, 256:768]
    got = colsum(view)
    torch.testing.assert_close(got, view.float().sum(0), rtol=1e-4, atol=2e-2)
    out = torch.ones(512, device="cuda")
    colsum(view, out)
    torch.testing.assert_close(out, 1 + view.float().sum(0), rtol=1e-4, atol=2e-2)


@pytest.mark.parametrize("p", [0.0, 0.1])
def test_fused_ffn_matches_torch(p):
    """functional.feed_forward (GEMM+bias+ReLU -> dropout kernel -> GEMM; fused ReLU/dropout backward) against
    plain torch.  With dropout the realised mask is read back through an identity block planted in linear2, so the
    torch reference applies exactly the same mask (and the keep rate / scale are checked)."""
    import torch
    import torch.nn.functional as F

    from pointcloudmatters_b200 import functional as PF

    torch.manual_seed(11)
    rows, E, Hd = 1000, 64, 32
    l1, l2 = torch.nn.Linear(E, Hd).cuda(), torch.nn.Linear(Hd, E).cuda()
    with torch.no_grad():
        l2.weight[:Hd] = torch.eye(Hd, device="cuda")
        l2.bias[:Hd] = 0
    x = torch.randn(rows, E, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    y = PF.feed_forward(xa, l1, l2, p, True)
    dy = torch.randn_like(y)
    y.backward(dy)
    got = (y.detach(), xa.grad.clone(), l1.weight.grad.clone(), l1.bias.grad.clone(), l2.weight.grad.clone(), l2.bias.grad.clone())
    for m in (l1, l2):
        m.zero_grad(set_to_none=True)
    h = F.relu(F.linear(xb, l1.weight, l1.bias))
    if p > 0:
        hd_seen = y.detach()[:, :Hd]                      # = dropped hidden (identity block, zero bias)
        keep = (hd_seen != 0) | (h.detach() <= 0)
        live = h.detach() > 1e-3
        rate = float((hd_seen[live] != 0).float().mean())
        assert abs(rate - (1 - p)) < 0.02, rate
        scale = float((hd_seen[live & keep] / h.detach()[live & keep]).median())
        assert abs(scale - 1 / (1 - p)) < 0.02, scale
        h = h * keep * scale
    yr = F.linear(h, l2.weight, l2.bias)
    yr.backward(dy)
    want = (yr.detach(), xb.grad, l1.weight.grad, l1.bias.grad, l2.weight.grad, l2.bias.grad)
    # y / dW2 / db2 see only operand rounding (<= 2e-2).  dx / dW1 / db1 pass through the ReLU gate, which the
    # product evaluates on ITS hidden activation (bf16 operands): ~0.1 % of the units sit close enough to zero to
    # flip against the fp32 reference, each contributing a whole gradient element (measured 3.5-5e-2 relative).
    for a, b, name in zip(got, want, ("y", "dx", "dW1", "db1", "dW2", "db2")):
        rel = float((a - b).norm() / b.norm())
        assert rel < (8e-2 if name in ("dx", "dW1", "db1") else 2e-2), (name, rel)


def test_two_layernorms_chain_splits_the_residual_gradient():
    """y1 = LN1(res + x) feeds BOTH a sub-block (here: a fixed linear map) and the residual of LN2.  The product
    hands LN1's backward the two gradient contributions separately (summed inside the kernel, no autograd add);
    result must equal plain torch, and the residual handle must alias y1's storage."""
    from pointcloudmatters_b200 import functional as PF

    torch.manual_seed(5)
    L, B, C = 37, 3, 256
    n1, n2 = torch.nn.LayerNorm(C).cuda(), torch.nn.LayerNorm(C).cuda()
    with torch.no_grad():
        for n in (n1, n2):
            n.weight.add_(0.2 * torch.randn(C, device="cuda")); n.bias.add_(0.2 * torch.randn(C, device="cuda"))
    W = torch.randn(C, C, device="cuda") / C ** 0.5
    x, res = torch.randn(L, B, C, device="cuda"), torch.randn(L, B, C, device="cuda")
    dy = torch.randn(L, B, C, device="cuda")

    def run(ln):
        xs, rs = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
        for n in (n1, n2):
            n.zero_grad(set_to_none=True)
        y1 = ln(xs, rs, n1)
        y2 = ln(y1 @ W, y1, n2)      # y1: sub-block input AND residual
        (y2 * dy).sum().backward()
        return y2.detach(), xs.grad, rs.grad, n1.weight.grad.clone(), n1.bias.grad.clone(), n2.weight.grad.clone(), y1

    ref = run(lambda a, r, n: F.layer_norm(r + a, (C,), n.weight, n.bias, n.eps))
    got = run(lambda a, r, n: PF.add_dropout_layernorm(a, r, n, 0.0, True))
    assert got[6]._pcm_res.data_ptr() == got[6].data_ptr()
    for a, b, name in zip(got[:6], ref[:6], ("y2", "dx", "dres", "dg1", "db1", "dg2")):
        assert float((a - b).norm() / b.norm()) < 1e-4, name


@pytest.mark.parametrize("R,C,training,relu", [(5000, 64, True, True), (777, 512, True, True), (300, 40, True, False),
                                               (1000, 128, False, True), (65536, 64, True, True)])
def test_fused_batchnorm_relu_matches_torch(R, C, training, relu):
    """csrc/batchnorm.cu (statistics + apply forward, reduce + apply backward) against F.batch_norm + F.relu:
    outputs, running statistics, dgamma / dbeta / dx; fp32 kernels with fp64 column sums -> 2e-5 relative."""
    from pointcloudmatters_b200 import functional as PF

    torch.manual_seed(R + C)
    bn_a, bn_b = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda(), torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).cuda()
    with torch.no_grad():
        w, b = 1 + 0.3 * torch.randn(C, device="cuda"), 0.3 * torch.randn(C, device="cuda")
        rm, rv = 0.2 * torch.randn(C, device="cuda"), 0.5 + torch.rand(C, device="cuda")
        for bn in (bn_a, bn_b):
            bn.weight.copy_(w); bn.bias.copy_(b); bn.running_mean.copy_(rm); bn.running_var.copy_(rv)
            bn.train(training)
    x = torch.randn(R, C, device="cuda") * 1.7 + 0.4
    dy = torch.randn(R, C, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = F.batch_norm(xa, bn_a.running_mean, bn_a.running_var, bn_a.weight, bn_a.bias, training, 0.01, 1e-3)
    ya = F.relu(ya) if relu else ya
    ya.backward(dy)
    yb = PF.batchnorm_relu(xb, bn_b, relu=relu)
    yb.backward(dy)
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-12))
    assert rel(yb, ya) < 2e-5
    assert rel(yb._pcm_bf16.float(), ya) < 1e-2
    assert rel(xb.grad, xa.grad) < 1e-4
    assert rel(bn_b.weight.grad, bn_a.weight.grad) < 1e-4 and rel(bn_b.bias.grad, bn_a.bias.grad) < 1e-4
    assert rel(bn_b.running_mean, bn_a.running_mean) < 1e-5 and rel(bn_b.running_var, bn_a.running_var) < 1e-5


@pytest.mark.parametrize("rows,E,p", [(6400, 512, 0.0), (1037, 512, 0.1), (333, 128, 0.1), (50, 256, 0.0)])
def test_fused_ffn32_equals_the_three_launch_path(rows, E, p):
    """csrc/ffn_fused.cu (one kernel each way) against the GEMM -> dropout -> GEMM composition of the same operator on the
    same bf16 operands and the SAME dropout mask (both derive it from the counter RNG with the same seed): forward,
    input gradient and all four parameter gradients."""
    from pointcloudmatters_b200 import functional as PF

    torch.manual_seed(5)
    l1, l2 = torch.nn.Linear(E, 32).cuda(), torch.nn.Linear(32, E).cuda()
    x = torch.randn(rows, E, device="cuda")
    dy = torch.randn(rows, E, device="cuda")
    res = []
    for fused in (True, False):
        PF._NO_FUSED_FFN = not fused
        PF.DROPOUT_RNG.offset = 0
        for m in (l1, l2):
            m.zero_grad(set_to_none=True)
        xa = x.clone().requires_grad_(True)
        y = PF.feed_forward(xa, l1, l2, p, True)
        y.backward(dy)
        res.append((y.detach().clone(), xa.grad.clone(), l1.weight.grad.clone(), l1.bias.grad.clone(), l2.weight.grad.clone(),
                    l2.bias.grad.clone()))
    PF._NO_FUSED_FFN = False
    for a, b, name in zip(res[0], res[1], ("y", "dx", "dW1", "db1", "dW2", "db2")):
        rel = float((a - b).norm() / b.norm())
        # the composition rounds the hidden activation to bf16 BEFORE the dropout scale, the fused kernel after: 2^-9 relative
        assert rel < 1e-2, (name, rel)
    kept = (res[0][0] != 0).float().mean()
    assert float(kept) > 0.5
