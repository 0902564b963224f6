"""Low-level Python bindings of the dense kernels of libpcm_b200.so (one function per C-ABI entry
point; tensors in, tensors out, launches on torch's current stream).  `functional.py` builds the
autograd-aware operators on top of these."""
from __future__ import annotations

import torch

from ._lib import check, current_stream, lib, ptr, require_cuda


class _Timer:
    """Optional live timing of one kernel family with CUDA events on the launching stream
    (bench.py's `roofline`): records (events, algorithmic flops, algorithmic bytes) per launch."""

    def __init__(self):
        self.enabled = False
        self.records = {}
        self.pair_overhead_ms = 0.0

    def start(self):
        self.enabled, self.records = True, {}
        self.pair_overhead_ms = self._calibrate()

    @staticmethod
    def _calibrate(n=200):
        """Elapsed time an event pair reports with NOTHING between the two records (median of n, queued behind
        a device-side spin so the host is ahead of the GPU exactly as in the measured steps).  Each timed launch
        carries this fixed cost of the second event's timestamp; `summary()` reports raw and corrected times."""
        torch.cuda._sleep(int(0.01 * 1.9e9))
        pairs = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); b.record()
            pairs.append((a, b))
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in pairs)
        return t[len(t) // 2]

    def stop(self):
        self.enabled = False

    def begin(self):
        if not self.enabled:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, e0, name, flops, nbytes, tag=None):
        if e0 is None:
            return
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.setdefault(name, []).append((e0, e1, flops, nbytes, tag))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, recs in self.records.items():
            ms = [r[0].elapsed_time(r[1]) for r in recs]
            shapes = {}
            for r, t in zip(recs, ms):
                d = shapes.setdefault(r[4], [0, 0.0, 0.0])
                d[0] += 1; d[1] += t; d[2] += r[2]
            raw = sum(ms)
            ms = [max(t - self.pair_overhead_ms, 0.0) for t in ms]
            out[name] = {"launches": len(recs), "total_ms": sum(ms), "avg_ms": sum(ms) / len(ms), "total_ms_raw_events": raw,
                         "event_pair_overhead_us": self.pair_overhead_ms * 1e3,
                         "flops": float(sum(r[2] for r in recs)), "bytes": float(sum(r[3] for r in recs)),
                         "by_shape": {str(k): {"launches": v[0], "total_ms": v[1], "tflops": v[2] / max(v[1], 1e-9) / 1e9}
                                      for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][1])}}
        return out


TIMER = _Timer()


def gemm_bf16(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.float32, bias=None, relu=False,
              accumulate=False, split_k=1):
    """C[m,n] (+)= sum_k A(m,k) B(n,k) (+bias) (ReLU) on tcgen05 (pcm_gemm_bf16).

    a: bf16, (M, K) if not a_mn else (K, M); b: bf16, (N, K) if not b_mn else (K, N); last dim
    contiguous.  Returns C (M, N) of out_dtype (fp32 or bf16); with `accumulate` adds into `out`."""
    require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    assert out.shape == (M, N) and out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == N
    t0 = TIMER.begin()
    check(lib.pcm_gemm_bf16(M, N, K, ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn), ptr(out),
                            out.stride(0), int(out.dtype == torch.bfloat16), ptr(bias), int(relu), int(accumulate),
                            int(split_k), current_stream()), "pcm_gemm_bf16")
    TIMER.end(t0, "gemm_tcgen05", 2.0 * M * N * K, 2.0 * (M * K + N * K) + M * N * out.element_size(),
              (M, N, K, 1, int(a_mn), int(b_mn), str(out.dtype)[6:], int(split_k)))
    return out


def gemm_ex(M, N, Kd, batch, a, a_mn, a_batch_rows, b, b_mn, b_batch_rows, out, *, c_mode=0, c_batch_rows=0,
            hs=(0, 0, 0), ldc=None, alpha=1.0, bias=None, relu=False, accumulate=False, split_k=1, a2=None, a2_from_col=0,
            hs_parts=(0, 0)):
    """Batched tcgen05 GEMM with head-split / head-merge output addressing (pcm_gemm_bf16_ex2).
    a, b: 2-D bf16 tensors spanning all batches (last dim contiguous); out: destination tensor.
    `a2` / `a2_from_col`: second A operand for the output columns >= a2_from_col; `hs_parts` = (columns per part,
    elements between the parts' tensors) for a head-split output that spans several (B, nh, L, 64) tensors."""
    require_cuda(a, b, out)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    if a2 is not None:
        assert a2.dtype == torch.bfloat16 and a2.shape == a.shape and a2.stride(1) == 1
    if ldc is None:
        ldc = out.stride(-2) if out.dim() >= 2 else out.shape[-1]
    t0 = TIMER.begin()
    check(lib.pcm_gemm_bf16_ex2(M, N, Kd, batch, ptr(a), a.stride(0), int(a_mn), a.shape[0], a_batch_rows,
                                ptr(b), b.stride(0), int(b_mn), b.shape[0], b_batch_rows, ptr(out), ldc,
                                int(out.dtype == torch.bfloat16), c_mode, c_batch_rows, hs[0], hs[1], hs[2],
                                float(alpha), ptr(bias), int(relu), int(accumulate), int(split_k), ptr(a2),
                                a2.stride(0) if a2 is not None else 0, int(a2_from_col), int(hs_parts[0]), int(hs_parts[1]),
                                current_stream()), "pcm_gemm_bf16_ex2")
    TIMER.end(t0, "gemm_tcgen05", 2.0 * M * N * Kd * batch,
              batch * (2.0 * (M * Kd + N * Kd) + M * N * out.element_size()),
              (M, N, Kd, batch, int(a_mn), int(b_mn), str(out.dtype)[6:], int(split_k)))
    return out


def gemm_dw_grouped(problems):
    """Run a list of weight-gradient products  out (M, N) fp32 += a^T b  -- a (K, M) bf16, b (K, N) bf16, last dim
    contiguous -- as ONE grouped tcgen05 launch per tile-width class (pcm_gemm_dw_grouped)."""
    import ctypes

    n = len(problems)
    if n == 0:
        return
    PA, PI = ctypes.c_void_p * n, ctypes.c_int * n
    A, B, C = PA(), PA(), PA()
    lda, ldb, ldc, M, N, Kd = PI(), PI(), PI(), PI(), PI(), PI()
    flops = nbytes = 0.0
    for i, (a, b, out) in enumerate(problems):
        assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and out.dtype == torch.float32
        assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1 and a.shape[0] == b.shape[0]
        assert out.shape == (a.shape[1], b.shape[1])
        A[i], B[i], C[i] = a.data_ptr(), b.data_ptr(), out.data_ptr()
        lda[i], ldb[i], ldc[i] = a.stride(0), b.stride(0), out.stride(0)
        M[i], N[i], Kd[i] = a.shape[1], b.shape[1], a.shape[0]
        flops += 2.0 * a.shape[1] * b.shape[1] * a.shape[0]
        nbytes += 2.0 * a.shape[0] * (a.shape[1] + b.shape[1]) + 4.0 * a.shape[1] * b.shape[1]
    t0 = TIMER.begin()
    check(lib.pcm_gemm_dw_grouped(n, A, lda, B, ldb, C, ldc, M, N, Kd, current_stream()), "pcm_gemm_dw_grouped")
    TIMER.end(t0, "gemm_tcgen05", flops, nbytes, ("grouped_dw", n, int(flops)))


def attn_softmax_fwd(S, Y, Zd, Z, L, Lp, Sk, Sp, nh, kpm, scale, p_drop, seed_base, seed_offset):
    check(lib.pcm_attn_softmax_fwd(Z, L, Lp, Sk, Sp, nh, ptr(S), ptr(kpm), float(scale), float(p_drop), ptr(seed_base),
                                   int(seed_offset), ptr(Y), ptr(Zd), current_stream()), "pcm_attn_softmax_fwd")


def attn_softmax_bwd(Y, dZ, Z, L, Lp, Sk, Sp, scale, p_drop, seed_base, seed_offset):
    check(lib.pcm_attn_softmax_bwd(Z, L, Lp, Sk, Sp, ptr(Y), ptr(dZ), float(scale), float(p_drop), ptr(seed_base),
                                   int(seed_offset), current_stream()), "pcm_attn_softmax_bwd")


def flash_attn_fwd(Qh, Kh, Vh, B, nh, L, S, kpm, scale, p_drop, seed_base, seed_offset, out=None):
    """Fused attention forward (pcm_flash_attn_fwd): head-split bf16 Q (B*nh*L, 64), K / V
    (B*nh*S, 64) -> token-major bf16 O (L*B, nh*64) and the log2-domain log-sum-exp (B*nh, L)."""
    require_cuda(Qh, Kh, Vh)
    E = nh * 64
    if out is None:
        out = torch.empty((L * B, E), dtype=torch.bfloat16, device=Qh.device)
    lse = torch.empty((B * nh, L), dtype=torch.float32, device=Qh.device)
    t0 = TIMER.begin()
    check(lib.pcm_flash_attn_fwd(B, nh, L, S, ptr(Qh), ptr(Kh), ptr(Vh), ptr(kpm), float(scale), float(p_drop),
                                 ptr(seed_base), int(seed_offset), ptr(out), out.stride(0), ptr(lse), current_stream()),
          "pcm_flash_attn_fwd")
    Z = B * nh
    TIMER.end(t0, "flash_attn_fwd", 4.0 * Z * L * S * 64, 2.0 * Z * (2 * L + 2 * S) * 64, (Z, L, S))
    return out, lse


def flash_attn_bwd(Qh, Kh, Vh, O_tok, dOh, lse, B, nh, L, S, kpm, scale, p_drop, seed_base, seed_offset, dQ, dK, dV):
    """Fused attention backward (pcm_flash_attn_bwd); dQ / dK / dV are token-major bf16 destinations
    (possibly column slices of a wider buffer: the row pitch is taken from their stride)."""
    require_cuda(Qh, Kh, Vh, O_tok, dOh, lse, dQ, dK, dV)
    Z = B * nh
    delta = torch.empty((Z, L), dtype=torch.float32, device=Qh.device)
    dq_acc = torch.empty((Z * L, 64), dtype=torch.float32, device=Qh.device)
    assert dK.stride(0) == dV.stride(0)
    t0 = TIMER.begin()
    check(lib.pcm_flash_attn_bwd(B, nh, L, S, ptr(Qh), ptr(Kh), ptr(Vh), ptr(O_tok), O_tok.stride(0), ptr(dOh), ptr(lse),
                                 ptr(kpm), float(scale), float(p_drop), ptr(seed_base), int(seed_offset), ptr(delta),
                                 ptr(dq_acc), ptr(dQ), dQ.stride(0), ptr(dK), ptr(dV), dK.stride(0), current_stream()),
          "pcm_flash_attn_bwd")
    TIMER.end(t0, "flash_attn_bwd", 10.0 * Z * L * S * 64, 2.0 * Z * (4 * L + 4 * S) * 64 + 8.0 * Z * L * 64, (Z, L, S))


def clip_adamw_step(param, grad, exp_avg, exp_avg_sq, hyper, sumsq, norm_out, param_bf16=None, zero_grad=False):
    """Fused clip + AdamW over flat fp32 buffers (pcm_clip_adamw_step_ex); hyper is a 9-float device
    tensor; `param_bf16` (optional) receives the bf16 copy of the updated parameters; `zero_grad`: leave the
    gradient buffer zeroed instead of holding the clipped gradient."""
    require_cuda(param, grad, exp_avg, exp_avg_sq, hyper, sumsq)
    check(lib.pcm_clip_adamw_step_ex(param.numel(), ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(hyper),
                                     ptr(sumsq), ptr(norm_out), ptr(param_bf16), int(zero_grad), current_stream()),
          "pcm_clip_adamw_step_ex")


def add_dropout_ln_fwd(x, res, gamma, beta, eps, p_drop, seed_base, seed_offset, want_bf16=False, pos=None, pos_row_div=1):
    """y = LayerNorm(res + dropout(x)); optional extra outputs yb = bf16(y) and ypb = bf16(y + pos[r // pos_row_div])
    (operands of the next sub-block's GEMMs).  Returns (y, yb, h, mean, rstd, ypb)."""
    rows, C = res.shape
    y = torch.empty_like(res)
    h = torch.empty_like(res)
    mean = torch.empty(rows, dtype=torch.float32, device=res.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=res.device)
    yb = torch.empty(res.shape, dtype=torch.bfloat16, device=res.device) if want_bf16 else None
    ypb = torch.empty(res.shape, dtype=torch.bfloat16, device=res.device) if pos is not None else None
    check(lib.pcm_add_dropout_ln_fwd_ex(rows, C, ptr(x), ptr(res), ptr(gamma), ptr(beta), float(eps), float(p_drop),
                                        ptr(seed_base), int(seed_offset), ptr(y), ptr(yb), ptr(h), ptr(mean), ptr(rstd),
                                        ptr(pos), int(pos_row_div), ptr(ypb), current_stream()), "pcm_add_dropout_ln_fwd_ex")
    return y, yb, h, mean, rstd, ypb


def add_dropout_ln_bwd(dy, h, mean, rstd, gamma, p_drop, seed_base, seed_offset, need_dx, dgamma=None, dbeta=None,
                       want_dx_bf16=False, dy_b=None, dx_colsum=None, dx_fp32=True):
    """dgamma / dbeta, when given, are ACCUMULATED into (the kernel adds with atomics); so is `dx_colsum` (C floats):
    the column sums of dx = the bias gradient of the linear layer that produced x.
    `dx_fp32=False` (needs want_dx_bf16): the fp32 dx is NOT written -- the returned tensor is an unwritten placeholder.
    Returns (dres, dx, dgamma, dbeta, dx_bf16)."""
    rows, C = h.shape
    dres = torch.empty_like(h)
    dx = (torch.empty_like(h) if (p_drop > 0 or not dx_fp32) else dres) if need_dx else None
    skip = need_dx and want_dx_bf16 and not dx_fp32
    if dgamma is None:
        dgamma = torch.zeros(C, dtype=torch.float32, device=h.device)
    if dbeta is None:
        dbeta = torch.zeros(C, dtype=torch.float32, device=h.device)
    dxb = torch.empty(h.shape, dtype=torch.bfloat16, device=h.device) if (want_dx_bf16 and need_dx) else None
    check(lib.pcm_add_dropout_ln_bwd_ex2(rows, C, ptr(dy), ptr(dy_b), ptr(h), ptr(mean), ptr(rstd), ptr(gamma), float(p_drop),
                                         ptr(seed_base), int(seed_offset), ptr(dres), ptr(None if skip else dx), ptr(dgamma), ptr(dbeta),
                                         ptr(dxb), ptr(dx_colsum if need_dx else None), current_stream()),
          "pcm_add_dropout_ln_bwd_ex2")
    return dres, dx, dgamma, dbeta, dxb


def colsum(src, out=None):
    """out[c] (+)= sum_r src[r, c] for a 2-D fp32 / bf16 tensor with contiguous columns."""
    rows, C = src.shape
    assert src.stride(1) == 1
    if out is None:
        out = torch.zeros(C, dtype=torch.float32, device=src.device)
    check(lib.pcm_colsum(rows, C, ptr(src), src.stride(0), int(src.dtype == torch.bfloat16), ptr(out), current_stream()),
          "pcm_colsum")
    return out


def colsum_grouped(problems):
    """[(src bf16 (rows, C), out fp32 (C))]: out += column sums of src, all problems in one launch (pcm_colsum_grouped)."""
    import ctypes

    n = len(problems)
    if n == 0:
        return
    PA, PI, PL = ctypes.c_void_p * n, ctypes.c_int * n, ctypes.c_longlong * n
    S, O, R, C, L = PA(), PA(), PL(), PI(), PL()
    for i, (src, out) in enumerate(problems):
        assert src.dtype == torch.bfloat16 and src.stride(1) == 1 and out.dtype == torch.float32 and out.is_contiguous()
        S[i], O[i], R[i], C[i], L[i] = src.data_ptr(), out.data_ptr(), src.shape[0], src.shape[1], src.stride(0)
    check(lib.pcm_colsum_grouped(n, S, R, C, L, O, current_stream()), "pcm_colsum_grouped")


def add_cast_bf16(a, b=None, b_row_div=1):
    """bf16(a + b) for token-major (rows, C) fp32 activations; b may be None or row-broadcast."""
    rows, C = a.shape
    out = torch.empty((rows, C), dtype=torch.bfloat16, device=a.device)
    check(lib.pcm_add_cast_bf16(rows, C, ptr(a), ptr(b), int(b_row_div), ptr(out), current_stream()), "pcm_add_cast_bf16")
    return out
