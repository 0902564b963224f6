"""tcgen05 GEMM (pcm_gemm_bf16) against a plain PyTorch fp32 reference of the same op on the same
bf16-rounded operands (products are exact in fp32, so only the accumulation order differs:
tolerance rtol 2e-4 / atol 2e-4 * sqrt(K/64); bf16 outputs additionally carry 2^-8 rounding)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops(M, N, K, a_mn, b_mn, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g).bfloat16()
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return a, b, A @ B.t()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 512), (1000, 512, 512), (128, 1024, 192),
                                   (6400, 512, 512), (520, 200, 136), (64, 64, 8)])
def test_gemm_layouts_and_tails(a_mn, b_mn, M, N, K):
    from pointcloudmatters_b200.kernels import gemm_bf16

    a, b, want = _ops(M, N, K, a_mn, b_mn)
    got = gemm_bf16(a, b, a_mn=a_mn, b_mn=b_mn)
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4 * math.sqrt(max(K, 64) / 64))


def test_gemm_bias_relu_bf16_out():
    from pointcloudmatters_b200.kernels import gemm_bf16

    a, b, want = _ops(777, 384, 512, False, False, seed=3)
    bias = torch.randn(384, device="cuda")
    got = gemm_bf16(a, b, bias=bias, relu=True)
    torch.testing.assert_close(got, torch.relu(want + bias), rtol=2e-4, atol=1e-3)
    got16 = gemm_bf16(a, b, bias=bias, out_dtype=torch.bfloat16)
    torch.testing.assert_close(got16.float(), (want + bias), rtol=1e-2, atol=1e-2)
    assert got16.dtype == torch.bfloat16


@pytest.mark.parametrize("split_k", [0, 1, 4, 37])
def test_gemm_split_k_accumulate(split_k):
    """dW = dY^T X shape: tiny output, very long K, accumulated atomically into an existing buffer."""
    from pointcloudmatters_b200.kernels import gemm_bf16

    M, N, K = 512, 512, 32960
    a, b, want = _ops(M, N, K, True, True, seed=5)
    base = torch.randn(M, N, device="cuda")
    out = base.clone()
    gemm_bf16(a, b, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=split_k)
    torch.testing.assert_close(out, base + want, rtol=1e-3, atol=2e-2)


def test_gemm_strided_views():
    """in_proj_weight slices / column slices of a wider activation matrix (pitch != width)."""
    from pointcloudmatters_b200.kernels import gemm_bf16

    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(300, 1024, device="cuda", generator=g).bfloat16()
    w = torch.randn(1536, 512, device="cuda", generator=g).bfloat16()
    xa = x[:, 512:]  # pitch 1024
    wk = w[512:1024]
    got = gemm_bf16(xa, wk)
    torch.testing.assert_close(got, xa.float() @ wk.float().t(), rtol=2e-4, atol=1e-3)
    out = torch.zeros(300, 2048, device="cuda")
    gemm_bf16(xa, wk, out=out[:, 1024:1536])
    torch.testing.assert_close(out[:, 1024:1536], xa.float() @ wk.float().t(), rtol=2e-4, atol=1e-3)
    assert float(out[:, :1024].abs().max()) == 0 and float(out[:, 1536:].abs().max()) == 0


def test_gemm_rejects_unaligned():
    from pointcloudmatters_b200._lib import PcmError
    from pointcloudmatters_b200.kernels import gemm_bf16

    a = torch.randn(64, 7, device="cuda").bfloat16()
    b = torch.randn(32, 7, device="cuda").bfloat16()
    with pytest.raises(PcmError):
        gemm_bf16(a, b)


@pytest.mark.parametrize("M,N,K,b_mn,bias", [(32, 128, 1280, False, True), (64, 2048, 10240, False, True),
                                            (64, 10240, 2048, True, False), (16, 408, 4096, True, False),
                                            (200, 520, 1032, False, False)])
def test_few_row_long_k_gemm_uses_k_split_and_matches(M, N, K, b_mn, bias):
    """functional._gemm_rows: activation GEMMs with < 74 output tiles and K >= 1024 (Diffusion-Policy denoiser)
    go through the launcher's K-split with fp32 atomics into a bias-initialised output; rows / columns that are
    not tile multiples must not be touched outside [M, N)."""
    import torch

    from pointcloudmatters_b200.functional import _gemm_rows

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K, N, device="cuda", generator=g) if b_mn else torch.randn(N, K, device="cuda", generator=g)).bfloat16()
    bv = torch.randn(N, device="cuda", generator=g) if bias else None
    got = _gemm_rows(a, w, b_mn=b_mn, bias=bv)
    want = a.float() @ (w.float() if b_mn else w.float().t())
    if bv is not None:
        want = want + bv
    assert got.shape == (M, N) and got.dtype == torch.float32
    assert float((got - want).norm() / want.norm()) < 2e-3


def test_grouped_weight_gradient_gemm_matches_individual_products():
    """pcm_gemm_dw_grouped: a mixed list of dW = dY^T X problems (decoder-size, encoder-size, FFN-32 both ways, ragged
    sizes, column-slice operands, more than 40 problems, two tile-width classes) accumulated into existing buffers."""
    from pointcloudmatters_b200.kernels import gemm_dw_grouped

    g = torch.Generator(device="cuda").manual_seed(11)
    shapes = [(512, 512, 6400)] * 9 + [(512, 512, 32960)] * 2 + [(32, 512, 6400), (512, 32, 6400), (1024, 512, 6528), (512, 1024, 700),
                                                                   (136, 200, 1000), (64, 16, 4096), (512, 512, 64), (96, 520, 333 * 8)]
    shapes = shapes + [(128, 64, 520)] * 30  # > 40 narrow problems: several launches of the 64-wide class
    probs, wants = [], []
    wide = torch.randn(6400, 1536, device="cuda", generator=g).bfloat16()  # column slices (pitch != width)
    for i, (M, N, K) in enumerate(shapes):
        if (M, N, K) == (512, 512, 6400) and i < 3:
            a = wide[:, i * 512:(i + 1) * 512]
        else:
            a = torch.randn(K, M, device="cuda", generator=g).bfloat16()
        b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
        base = torch.randn(M, N, device="cuda", generator=g)
        out = base.clone()
        probs.append((a, b, out))
        wants.append(base + a.float().t() @ b.float())
    gemm_dw_grouped(probs)
    torch.cuda.synchronize()
    for (a, b, out), want, (M, N, K) in zip(probs, wants, shapes):
        torch.testing.assert_close(out, want, rtol=1e-3, atol=2e-4 * math.sqrt(K / 64) * 4, msg=str((M, N, K)))
