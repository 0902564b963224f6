"""Operator layer of the Diffusion-Policy denoiser -- channel-last
Conv1d / ConvTranspose1d on the tcgen05 GEMM, fused GroupNorm+Mish(+FiLM+residual), Mish.

The reference runs `nn.Conv1d(k=5)` / `nn.ConvTranspose1d(4, 2, 1)` through cuDNN on (B, C, T)
tensors with T in {16, 8, 4} (conv1d_components.py:8-45).  Here activations are (B, T, C): every
convolution is one `pcm_gemm_bf16` over rows = B*T against the weight IN ITS TORCH LAYOUT viewed as a
matrix -- (Cout, Cin*k) K-major for Conv1d, (Cin, Cout*k) MN-major for ConvTranspose1d -- so the bf16
shadow of the flat parameter buffer is the GEMM operand and the weight-gradient GEMM accumulates
straight into the flat gradient.  `pcm_conv1d_unfold` / `pcm_conv1d_fold` (csrc/unet1d.cu) move
data between (B, T, C) and the (rows, C*k) tap-column matrices.  CUDA only, no fallback.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import kernels as K
from ._lib import check, current_stream, lib, ptr
from .functional import _dw, _gemm_rows, _grad_slot, _need_cuda, _wb


def _r8(n):
    return -(-n // 8) * 8


def _hint(x):
    """bf16 copy a producing kernel attached to activation `x` (groupnorm_mish), or None."""
    t = getattr(x, "_pcm_bf16", None)
    return t if (t is not None and t.shape == x.shape and t.device == x.device) else None


def _bf16_of(x):
    if x.dtype == torch.bfloat16:
        return x
    return K.add_cast_bf16(x.reshape(-1, x.shape[-1])).view(x.shape)


def _unfold(x, k, stride, pad, R):
    """x (B, L, C) fp32 | bf16 contiguous -> col (B*R, r8(C*k)) bf16, column c*k + tap."""
    B, L, C = x.shape
    ldc = _r8(C * k)
    col = torch.empty((B * R, ldc), dtype=torch.bfloat16, device=x.device)
    check(lib.pcm_conv1d_unfold(B, L, C, k, stride, pad, R, ptr(x), int(x.dtype == torch.bfloat16), C, ptr(col), ldc,
                                current_stream()), "pcm_conv1d_unfold")
    return col


def _fold(col, B, L, C, k, stride, pad, R, bias=None):
    """col (B*R, >= C*k) fp32 -> y (B, L, C) fp32 (+ bias)."""
    y = torch.empty((B, L, C), dtype=torch.float32, device=col.device)
    check(lib.pcm_conv1d_fold(B, L, C, k, stride, pad, R, ptr(col), col.stride(0), ptr(bias), ptr(y), None,
                              current_stream()), "pcm_conv1d_fold")
    return y


def _weight_matrix(weight, rows, cols, ldc, rows_p):
    """bf16 (rows_p, ldc) operand of a conv weight viewed (rows, cols); a view of the bf16 parameter
    shadow when no padding is needed (all layers but the 7-channel input / output convolutions)."""
    wm = _wb(weight).reshape(rows, cols)
    if ldc != cols or rows_p != rows:
        wm = F.pad(wm, (0, ldc - cols, 0, rows_p - rows))
    return wm


def _accumulate_dw(weight, rows, cols, a, b, pad_rows, pad_cols):
    """dW (rows, cols) += a^T b with a (M, rows_p), b (M, cols_p) bf16.  Accumulates into the parameter's own
    gradient buffer when there is one (returns None), else returns a fresh gradient in weight's shape."""
    slot = _grad_slot(weight)
    if not pad_rows and not pad_cols:
        dw = slot.view(rows, cols) if slot is not None else torch.zeros((rows, cols), dtype=torch.float32, device=a.device)
        _dw(a, b, dw)  # queued with the step's other weight gradients when the destination is the flat gradient
        return None if slot is not None else dw.view(weight.shape)
    tmp = torch.zeros((a.shape[1], b.shape[1]), dtype=torch.float32, device=a.device)
    K.gemm_bf16(a, b, a_mn=True, b_mn=True, out=tmp, accumulate=True, split_k=0)
    dw = tmp[:rows, :cols]
    if slot is not None:
        slot.view(rows, cols).add_(dw)
        return None
    return dw.contiguous().view(weight.shape)


def _accumulate_db(bias, src):
    slot = _grad_slot(bias)
    db = K.colsum(src, slot)
    return None if slot is not None else db


class _Conv1dCL(torch.autograd.Function):
    """nn.Conv1d on channel-last activations: y (B, R, Cout) = unfold(x) @ W(Cout, Cin*k)^T + b."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, xb):
        B, L, Cin = x.shape
        Cout, _, k = weight.shape
        R = (L + 2 * pad - k) // stride + 1
        Kc = Cin * k
        ldc, Np = _r8(Kc), _r8(Cout)
        pointwise = k == 1 and stride == 1 and pad == 0 and ldc == Kc
        if pointwise:
            col = (xb if xb is not None else _bf16_of(x)).reshape(B * L, Cin)
        else:
            col = _unfold(xb if xb is not None else x, k, stride, pad, R)
        wm = _weight_matrix(weight, Cout, Kc, ldc, Np)
        bp = bias if (bias is None or Np == Cout) else F.pad(bias, (0, Np - Cout))
        y = _gemm_rows(col, wm, bias=bp)
        ctx.geom = (B, L, Cin, Cout, k, stride, pad, R, ldc, Np, pointwise)
        ctx.params = (weight, bias)
        ctx.save_for_backward(col, wm)
        if Np != Cout:
            y = y[:, :Cout].contiguous()
        return y.view(B, R, Cout)

    @staticmethod
    def backward(ctx, dy):
        col, wm = ctx.saved_tensors
        B, L, Cin, Cout, k, stride, pad, R, ldc, Np, pointwise = ctx.geom
        weight, bias = ctx.params
        dy2 = dy.reshape(B * R, Cout)
        if Np != Cout:
            dy2 = F.pad(dy2, (0, Np - Cout))
        dyb = K.add_cast_bf16(dy2.contiguous())
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dcol = _gemm_rows(dyb, wm, b_mn=True)  # (M, Np) x W(Np, ldc) -> (M, ldc)
            dx = dcol.view(B, L, Cin) if pointwise else _fold(dcol, B, L, Cin, k, stride, pad, R)
        if ctx.needs_input_grad[1]:
            dw = _accumulate_dw(weight, Cout, Cin * k, dyb, col, Np != Cout, ldc != Cin * k)
        if bias is not None and ctx.needs_input_grad[2]:
            db = _accumulate_db(bias, dyb[:, :Cout] if Np != Cout else dyb)
        return dx, dw, db, None, None, None


class _ConvTranspose1dCL(torch.autograd.Function):
    """nn.ConvTranspose1d on channel-last activations: fold(x @ W(Cin, Cout*k)) + b."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, xb):
        B, L, Cin = x.shape
        _, Cout, k = weight.shape
        Lout = (L - 1) * stride - 2 * pad + k
        if Cin % 8 or (Cout * k) % 8:
            raise NotImplementedError("ConvTranspose1d channels must be multiples of 8 (U-Net widths are)")
        xb = (xb if xb is not None else _bf16_of(x)).reshape(B * L, Cin)
        wm = _wb(weight).reshape(Cin, Cout * k)
        ycol = _gemm_rows(xb, wm, b_mn=True)  # (M, Cin) x W(Cin, Cout*k)
        y = _fold(ycol, B, Lout, Cout, k, stride, pad, L, bias)
        ctx.geom = (B, L, Cin, Cout, k, stride, pad, Lout)
        ctx.params = (weight, bias)
        ctx.save_for_backward(xb, wm)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wm = ctx.saved_tensors
        B, L, Cin, Cout, k, stride, pad, Lout = ctx.geom
        weight, bias = ctx.params
        dy = dy.contiguous()
        dycol = _unfold(dy, k, stride, pad, L)  # (B*L, Cout*k): dycol[(b,l), co*k+tap] = dy[b, l*s+tap-p, co]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm_rows(dycol, wm).view(B, L, Cin)  # (M, Cout*k) x W(Cin, Cout*k)^T
        if ctx.needs_input_grad[1]:
            dw = _accumulate_dw(weight, Cin, Cout * k, xb, dycol, False, False)
        if bias is not None and ctx.needs_input_grad[2]:
            db = _accumulate_db(bias, dy.view(B * Lout, Cout))
        return dx, dw, db, None, None, None


def conv1d_cl(x, weight, bias=None, stride=1, padding=0):
    """x (B, L, Cin) channel-last, weight (Cout, Cin, k) [nn.Conv1d layout] -> (B, Lout, Cout) fp32."""
    _need_cuda(x)
    return _Conv1dCL.apply(x.contiguous(), weight, bias, int(stride), int(padding), _hint(x))


def conv_transpose1d_cl(x, weight, bias=None, stride=1, padding=0):
    """x (B, L, Cin) channel-last, weight (Cin, Cout, k) [nn.ConvTranspose1d layout] -> (B, Lout, Cout)."""
    _need_cuda(x)
    return _ConvTranspose1dCL.apply(x.contiguous(), weight, bias, int(stride), int(padding), _hint(x))


class _GroupNormMish(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, G, eps, film, res):
        B, T, C = x.shape
        dev = x.device
        y = torch.empty_like(x)
        yb = torch.empty(x.shape, dtype=torch.bfloat16, device=dev)
        mean = torch.empty(B * G, dtype=torch.float32, device=dev)
        rstd = torch.empty(B * G, dtype=torch.float32, device=dev)
        check(lib.pcm_groupnorm_mish_fwd(B, T, C, G, ptr(x), ptr(gamma), ptr(beta), float(eps), ptr(film), ptr(res),
                                         ptr(y), ptr(yb), ptr(mean), ptr(rstd), current_stream()), "pcm_groupnorm_mish_fwd")
        ctx.G = G
        ctx.params = (gamma, beta)
        ctx.has = (film is not None, res is not None)
        ctx.save_for_backward(x, mean, rstd, film)
        ctx.mark_non_differentiable(yb)
        return y, yb

    @staticmethod
    def backward(ctx, dy, _dyb=None):
        x, mean, rstd, film = ctx.saved_tensors
        gamma, beta = ctx.params
        B, T, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        sg, sb = _grad_slot(gamma), _grad_slot(beta)
        dg = sg if sg is not None else torch.zeros(C, dtype=torch.float32, device=x.device)
        db = sb if sb is not None else torch.zeros(C, dtype=torch.float32, device=x.device)
        dfilm = torch.empty_like(film) if ctx.has[0] else None
        check(lib.pcm_groupnorm_mish_bwd(B, T, C, ctx.G, ptr(x), ptr(gamma), ptr(beta), ptr(mean), ptr(rstd), ptr(film),
                                         ptr(dy), ptr(dx), ptr(dg), ptr(db), ptr(dfilm), current_stream()),
              "pcm_groupnorm_mish_bwd")
        return (dx, None if sg is not None else dg, None if sb is not None else db, None, None, dfilm,
                dy if ctx.has[1] else None)


def groupnorm_mish(x, gn, film=None, res=None):
    """film_scale * Mish(GroupNorm(x)) + film_bias (+ res) on channel-last x (B, T, C) fp32;
    film (B, 2C) = [scale | bias].  The returned tensor carries its bf16 copy (`_pcm_bf16`), the operand
    of the next convolution."""
    _need_cuda(x)
    y, yb = _GroupNormMish.apply(x.contiguous(), gn.weight, gn.bias, gn.num_groups, gn.eps,
                                 None if film is None else film.contiguous(), None if res is None else res.contiguous())
    y._pcm_bf16 = yb
    return y


class _Mish(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = torch.empty_like(x)
        check(lib.pcm_mish_fwd(x.numel(), ptr(x), ptr(y), None, current_stream()), "pcm_mish_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        check(lib.pcm_mish_bwd(x.numel(), ptr(x), ptr(dy.contiguous()), ptr(dx), current_stream()), "pcm_mish_bwd")
        return dx


def mish(x):
    _need_cuda(x)
    return _Mish.apply(x.contiguous().float())
