"""Operator layer of the B200 training step: every dense / fused op the policy modules use.

Each function is the single call site of one C-ABI kernel family of libpcm_b200.so (through an
autograd.Function).  Numerics policy (stated tolerance): GEMM / attention operands
are rounded to bf16 and accumulated in fp32 on the tensor cores; everything else (LayerNorm,
BatchNorm statistics, softmax statistics, losses, optimizer, master weights) is fp32.
CUDA only -- there is no CPU fallback; a CPU tensor raises.
"""
from __future__ import annotations

import math
import os
import weakref

import torch
import torch.nn.functional as F

from . import kernels as K
from ._lib import PcmError

COMPUTE_DTYPE = torch.bfloat16


def _need_cuda(t):
    if not t.is_cuda:
        raise PcmError("pointcloudmatters_b200 runs on CUDA tensors only (no CPU fallback)")


class _Bf16Shadow:
    """bf16 copies of trainers' flat fp32 parameter buffers (trainer.FlatState.param_bf16), kept in
    sync by the fused AdamW kernel: GEMM weight operands are views into them, so no per-use fp32 ->
    bf16 conversion kernels run inside the step.  Outside a trainer (unit tests, inference on a
    plain module) weights are converted on the fly -- same kernels, same numerics.
    Registrations are weak: an entry disappears with its FlatState, so a later allocation that
    reuses the freed address range can never alias a stale copy."""

    def __init__(self):
        self.entries = {}  # id -> (weakref to FlatState, base_ptr, nbytes)
        self.grad_ranges = {}  # id -> (grad base_ptr, nbytes) of the same FlatStates' flat gradient buffers
        self.inference_cache = {}  # base ptr -> (weakref(Parameter), version, bf16 copy); no_grad only

    def register(self, flat_state):
        import weakref

        key = id(flat_state)

        def _drop(_r, k=key):
            self.entries.pop(k, None)
            self.grad_ranges.pop(k, None)

        ref = weakref.ref(flat_state, _drop)
        self.entries[key] = (ref, flat_state.param.data_ptr(), flat_state.param.numel() * 4)
        self.grad_ranges[key] = (flat_state.grad.data_ptr(), flat_state.grad.numel() * 4)

    def owns_grad(self, g):
        """True when `g` lives inside the flat gradient buffer of a live trainer (FlatState)."""
        ptr_ = g.data_ptr()
        for base, nbytes in self.grad_ranges.values():
            if 0 <= ptr_ - base < nbytes:
                return True
        return False

    def view(self, w):
        if self.entries and w.dtype == torch.float32 and w.is_contiguous():
            ptr_ = w.data_ptr()
            for ref, base, nbytes in self.entries.values():
                off = ptr_ - base
                if 0 <= off < nbytes:
                    fs = ref()
                    if fs is not None and fs.param_bf16.device == w.device:
                        return fs.param_bf16[off // 4: off // 4 + w.numel()].view(w.shape)
        if not torch.is_grad_enabled() and w.dtype == torch.float32 and w.is_contiguous():
            # inference (no_grad) outside a trainer: convert each PARAMETER once, not once per call -- a sampling
            # loop calls the denoiser 100 times on the same 255 M parameters.  An entry is valid only while the very
            # Parameter object it was made from is alive, sits at the same address and has the same version counter
            # (torch bumps it on every in-place update, load_state_dict included), so neither a freed-and-reused
            # address nor changed weights can ever return a stale copy.  Views of a parameter (reshaped conv
            # weights) resolve to their base.
            base = w._base if w._base is not None else w
            if isinstance(base, torch.nn.Parameter) and base.is_contiguous():
                hit = self.inference_cache.get(base.data_ptr())
                if hit is None or hit[0]() is not base or hit[1] != base._version:
                    import weakref

                    hit = self.inference_cache[base.data_ptr()] = (weakref.ref(base), base._version,
                                                                   base.detach().to(torch.bfloat16))
                    if len(self.inference_cache) > 4096:  # drop entries of dead parameters
                        for k in [k for k, v in self.inference_cache.items() if v[0]() is None]:
                            del self.inference_cache[k]
                off = (w.data_ptr() - base.data_ptr()) // 4
                return hit[2].reshape(-1)[off: off + w.numel()].view(w.shape)
        return w.to(torch.bfloat16)


BF16_SHADOW = _Bf16Shadow()


def _wb(w):
    """bf16 operand form of an fp32 master weight."""
    return BF16_SHADOW.view(w)


def _grad_slot(p):
    """The trainer's flat-gradient view of a parameter: weight-gradient kernels accumulate straight into it and the
    backward returns None for that input, so autograd launches no zero-fill / add kernels.
    ONLY gradients that live inside a registered `trainer.FlatState` buffer are taken (address-range test, like the
    bf16 shadow): for any other `.grad` (torch DDP / Lightning with `accumulate_grad_batches`,
    `zero_grad(set_to_none=False)`, `gradient_as_bucket_view`) the gradient is RETURNED to autograd, so
    AccumulateGrad and the hooks registered on it (the DDP reducer's) run as usual."""
    g = p.grad if (p is not None and p.is_leaf) else None
    if (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape and g.device == p.device
            and BF16_SHADOW.owns_grad(g)):
        return g
    return None


# bf16 operand copies that a kernel produced as a by-product of the fp32 tensor it wrote:
#  * forward: attached to the fp32 activation as attributes (`_pcm_bf16`, `_pcm_bf16_pos`) by
#    add_dropout_layernorm(); the next projection GEMM picks them up instead of running a cast pass;
#  * backward: bf16(dx) of the LayerNorm backward, keyed by the gradient tensor's address and consumed
#    (popped) by the sub-block backward that receives that very tensor from autograd.  Entries hold
#    a reference to the fp32 gradient (its memory cannot be recycled under a live key) and the table
#    is emptied by the next LayerNorm backward, so a key can never outlive its tensor.
_GRAD_BF16 = {}
# fp32 LayerNorm-backward dx tensors that were NOT written (only their bf16 copy exists, see _AddDropoutLN.backward):
# address -> weak reference.  A consumer that misses the side-channel entry for one that is still alive (so the address
# cannot have been recycled) must not read it: _grad_bf16 raises instead of casting garbage.
_DX_UNWRITTEN = {}
_NO_LN_DX_BF16_ONLY = bool(int(os.environ.get("PCM_LN_DX_FP32", "0")))  # A/B switch: also write the fp32 dx
# second gradient tensor travelling with the one autograd carries (two-handle outputs, see _FillHeadRows): keyed by the
# carried gradient's address, value (second gradient, carried gradient); popped by the consumer
_GRAD_EXTRA = {}


def _act_bf16(x, rows, C):
    """bf16 (rows, C) copy attached to activation `x`, or None."""
    t = getattr(x, "_pcm_bf16", None)
    if t is not None and t.numel() == rows * C and t.device == x.device:
        return t.view(rows, C)
    return None


def _act_pos_bf16(x, pos, rows, C):
    """bf16(x + pos) attached to activation `x` for exactly this `pos` object, or None."""
    e = getattr(x, "_pcm_bf16_pos", None)
    if e is not None and e[0] is pos and e[1].numel() == rows * C:
        return e[1].view(rows, C)
    return None


def _grad_bf16(g, rows, C, bias=None):
    """bf16 copy of incoming gradient `g` left by the kernel that produced it, else a cast pass.
    With `bias` (the bias Parameter of the linear layer whose output gradient `g` is): returns (bf16 copy, done) where
    `done` says the producing kernel has already accumulated colsum(g) into that parameter's flat gradient."""
    e = _GRAD_BF16.pop(g.data_ptr(), None)
    if e is not None and e[1].data_ptr() == g.data_ptr() and e[1].numel() == g.numel() == rows * C and g.is_contiguous():
        out = e[0].view(rows, C)
        _DX_UNWRITTEN.pop(g.data_ptr(), None)
        return (out, len(e) > 2 and e[2] is bias and bias is not None) if bias is not None else out
    ref = _DX_UNWRITTEN.pop(g.data_ptr(), None)
    if ref is not None and ref() is not None:
        raise PcmError("gradient tensor holds no fp32 data (the LayerNorm backward wrote only its bf16 copy) and its "
                       "side-channel entry is gone: backward order violated the producer-before-next-LayerNorm assumption")
    g2 = g.reshape(rows, C)
    out = K.add_cast_bf16(g2 if g2.is_contiguous() else g2.contiguous())
    return (out, False) if bias is not None else out


_NO_ROW_SPLIT = bool(int(os.environ.get("PCM_NO_ROW_SPLIT", "0")))  # A/B switch for tools/


def _gemm_rows(a, b, *, b_mn=False, bias=None):
    """(M, K) x B -> (M, N) fp32 for activation GEMMs with few rows and a long reduction (the Diffusion-Policy
    denoiser: rows = B*T is 64 ... 2048 while K = Cin*k reaches 10 240): with one CTA per 128-wide output tile only a handful of SMs would stream the
    weights.  When the tile count is under half the SMs the launcher's K-split (split_k = 0: slices fill one
    wave, fp32 atomics into a bias-initialised output) is used instead."""
    M, Kd = a.shape
    N = b.shape[1] if b_mn else b.shape[0]
    if (-(-M // 128)) * (-(-N // 128)) <= 74 and Kd >= 1024 and not _NO_ROW_SPLIT:
        out = (bias.expand(M, N).contiguous() if bias is not None
               else torch.zeros((M, N), dtype=torch.float32, device=a.device))
        return K.gemm_bf16(a, b, b_mn=b_mn, out=out, accumulate=True, split_k=0)
    return K.gemm_bf16(a, b, b_mn=b_mn, bias=bias)


class _LinearTC(torch.autograd.Function):
    """y = x W^T (+b) (ReLU) on the tcgen05 GEMM; backward dX = dY W and dW = dY^T X read dY, W, X in
    place through the MN-major operand forms of the same kernel (no transposes)."""

    @staticmethod
    def forward(ctx, x2, weight, bias, relu, out_bf16, xb_hint=None):
        xb = xb_hint if xb_hint is not None else (x2 if x2.dtype == torch.bfloat16 else x2.to(torch.bfloat16))
        wb = _wb(weight)
        if relu or out_bf16:
            y = K.gemm_bf16(xb, wb, bias=bias, relu=relu, out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
        else:
            y = _gemm_rows(xb, wb, bias=bias)
        ctx.relu, ctx.has_bias, ctx.x_dtype = relu, bias is not None, x2.dtype
        ctx.params = (weight, bias)
        ctx.save_for_backward(xb, wb, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xb, wb, y = ctx.saved_tensors
        if ctx.relu:
            dy = dy * (y > 0)
            dyb = dy.contiguous() if dy.dtype == torch.bfloat16 else dy.to(torch.bfloat16)
        elif dy.dtype == torch.bfloat16:
            dyb = dy.contiguous()
        else:
            dyb = _grad_bf16(dy, dy.shape[0], dy.shape[1])
        M, N = dyb.shape
        Kin = xb.shape[1]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _gemm_rows(dyb, wb, b_mn=True)  # (M, N) x W(N, Kin) -> (M, Kin)
            if ctx.x_dtype == torch.bfloat16:
                dx = dx.to(torch.bfloat16)
        weight, bias = ctx.params
        if ctx.needs_input_grad[1]:
            slot = _grad_slot(weight)
            dw = slot if slot is not None else torch.zeros((N, Kin), dtype=torch.float32, device=dyb.device)
            _dw(dyb, xb, dw)
            if slot is not None:
                dw = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            slot = _grad_slot(bias)
            db = K.colsum(dyb, slot)
            if slot is not None:
                db = None
        return dx, dw, db, None, None, None


def linear(x, weight, bias=None, relu=False, out_bf16=False):
    """y = x @ weight^T (+ bias) (+ ReLU); x (..., K) fp32 or bf16, weight (N, K) fp32 master.

    Tensor-core path (pcm_gemm_bf16, tcgen05) whenever TMA's alignment rules allow (K, N multiples
    of 8; a narrow K over many rows is zero-padded to 16); the handful of tiny embedding / head
    linears with K or N in {1, 3, 7, 9} on B or B*Q rows are plain fp32 library GEMMs (well under 0.1%
    of the step's FLOPs)."""
    _need_cuda(x)
    N, Kin = weight.shape
    lead = x.shape[:-1]
    if Kin % 8 == 0 and N % 8 == 0 and Kin >= 16:
        x2 = x.reshape(-1, Kin)
        if x2.stride(-1) != 1 or (x2.stride(0) % 8) or (x2.data_ptr() % 16):
            x2 = x2.contiguous()
        if weight.stride(-1) != 1:
            weight = weight.contiguous()
        y = _LinearTC.apply(x2, weight, bias, relu, out_bf16, _act_bf16(x, x2.shape[0], Kin))
        return y.view(*lead, N)
    if Kin < 16 and N % 8 == 0 and x.numel() // max(Kin, 1) >= 4096 and x.dtype == torch.float32:
        # narrow input layer over many rows (PointNet conv1: 6 channels x 65 536 points): zero-pad K
        # to one UMMA step and use the tensor-core path instead of an fp32 SIMT library GEMM
        return linear(F.pad(x, (0, 16 - Kin)), F.pad(weight, (0, 16 - Kin)), bias, relu, out_bf16)
    y = F.linear(x.float(), weight, bias)
    return F.relu(y) if relu else y


def attention(q, k, v, key_padding_mask=None, dropout_p=0.0, training=False):
    """softmax(q k^T / sqrt(d) + mask) v on (B, h, L, d) / (B, h, S, d) tensors -> (B, h, L, d) fp32."""
    _need_cuda(q)
    d = q.shape[-1]
    scores = (q.to(COMPUTE_DTYPE) @ k.to(COMPUTE_DTYPE).transpose(-1, -2)).float() * (1.0 / math.sqrt(d))
    if key_padding_mask is not None:
        B, S = key_padding_mask.shape
        scores = scores.masked_fill(key_padding_mask.view(B, 1, 1, S), float("-inf"))
    attn = torch.softmax(scores, dim=-1)
    if training and dropout_p > 0:
        attn = F.dropout(attn, dropout_p, True)
    return (attn.to(COMPUTE_DTYPE) @ v.to(COMPUTE_DTYPE)).float()


_WORKSPACE = {}


class _DropoutRNG:
    """Seeds for the in-kernel dropout masks without host synchronisation: a device-resident
    64-bit base (advanced once per training step by the trainer -- also under CUDA-graph replay)
    plus a per-call-site offset that restarts at every step."""

    def __init__(self):
        self.base = {}
        self.offset = 0

    def base_for(self, device):
        key = str(device)
        if key not in self.base:
            self.base[key] = torch.randint(1, 2 ** 62, (1,), dtype=torch.int64).to(device)
        return self.base[key]

    def next_offset(self):
        self.offset += 1
        return self.offset * 0x100000001B3

    def new_step(self):
        """Call once per training step (outside any graph capture the increment is a tiny kernel)."""
        for t in self.base.values():
            t.add_(1)
        self.offset = 0
        ZERO_ARENA.reset()


DROPOUT_RNG = _DropoutRNG()


class _ZeroArena:
    """Small zero-initialised accumulators of a step (BatchNorm / set-abstraction statistics, loss sums, scatter counters:
    ~20 per step, each a separate fill kernel as `torch.zeros`) are carved out of ONE zero-filled buffer per stream and
    step: one fill instead of twenty.  A region is handed out once, so it is zero when its user first touches it.  The
    buffer is keyed by (device, stream) -- it is filled on the stream that uses it -- and is dropped at every step
    boundary and whenever graph capture starts or ends (a buffer filled outside a capture must not be consumed inside:
    the replay would find it dirty)."""

    CAP = 8 << 20

    def __init__(self):
        self.bufs = {}

    def reset(self):
        self.bufs.clear()

    def zeros(self, shape, dtype, device):
        device = torch.device(device)
        n = int(math.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        if device.type != "cuda" or n > self.CAP // 4 or n == 0:
            return torch.zeros(shape, dtype=dtype, device=device)
        capturing = torch.cuda.is_current_stream_capturing()
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        ent = self.bufs.get(key)
        need = (n + 255) & ~255
        if ent is None or ent[1] + need > self.CAP or ent[2] != capturing:
            ent = [torch.zeros(self.CAP, dtype=torch.uint8, device=device), 0, capturing]
            self.bufs[key] = ent
        v = ent[0][ent[1]: ent[1] + n].view(dtype).view(shape)
        ent[1] += need
        return v


ZERO_ARENA = _ZeroArena()


def _zeros(shape, dtype, device):
    return ZERO_ARENA.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype, device)


def _workspace(name, shape, dtype, device, zero=False):
    """Cached scratch tensor.  `zero=True` buffers are zero-filled once; their users only ever write
    the valid (unpadded) region, so the padding stays zero across reuses."""
    key = (name, tuple(shape), dtype, str(device))
    t = _WORKSPACE.get(key)
    if t is None:
        t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
        _WORKSPACE[key] = t
    return t


def _ceil64(x):
    return (x + 63) // 64 * 64


def _tok_bf16(x, pos, L, B, E):
    """bf16 token-major (L*B, E) operand of a projection GEMM: x (+ pos), pos (L, B, E) or (L, 1, E)."""
    x2 = x.reshape(L * B, E)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    if pos is None:
        return K.add_cast_bf16(x2)
    if pos.shape[1] == B:
        p2 = pos.reshape(L * B, E)
        return K.add_cast_bf16(x2, p2 if p2.is_contiguous() else p2.contiguous())
    return K.add_cast_bf16(x2, pos.reshape(L, E).contiguous(), b_row_div=B)


def _attn_core_fwd(Qh, Kh, Vh, L, S, B, nh, kpm, p_drop, dev):
    """dropout(softmax(Q K^T / 8 + mask)) V for head_dim 64 in ONE fused tcgen05 kernel
    (csrc/flash_attn.cu): returns the token-major bf16 output and what backward needs (the
    log-sum-exp per row; probabilities are recomputed, never stored)."""
    seed_base = DROPOUT_RNG.base_for(dev) if p_drop > 0 else None
    seed = DROPOUT_RNG.next_offset() if p_drop > 0 else 0
    kpm_u8 = kpm.to(torch.uint8).contiguous() if kpm is not None else None
    O_tok, lse = K.flash_attn_fwd(Qh, Kh, Vh, B, nh, L, S, kpm_u8, 0.125, p_drop, seed_base, seed)
    return O_tok, lse, kpm_u8, (seed_base, seed)


def _attn_core_bwd(dOh, Qh, Kh, Vh, O_tok, lse, kpm_u8, L, S, B, nh, aux, p_drop, dQ_out, dK_out, dV_out):
    """Backward of _attn_core_fwd: writes token-major bf16 dQ / dK / dV into the given (possibly
    column-sliced) destinations."""
    seed_base, seed = aux
    K.flash_attn_bwd(Qh, Kh, Vh, O_tok, dOh, lse, B, nh, L, S, kpm_u8, 0.125, p_drop, seed_base, seed,
                     dQ_out, dK_out, dV_out)


# Weight-gradient products whose destination is a trainer's flat gradient are not on the backward's critical path (nothing
# downstream reads them before the gradient exchange): while a queue is installed (trainer, during backward) they are
# collected and run as ONE grouped tcgen05 launch per flush (kernels.gemm_dw_grouped) instead of ~110 small launches.
DW_QUEUE = None  # None = launch immediately; list = [(stream id, dtok, xb, out)]


def _dw(dtok, xb, out):
    """out (N, K) += dtok^T xb -- weight gradient in the tensors' own layouts (MN-major operands);
    tile width and K split are chosen by the launcher (split_k = 0)."""
    if (DW_QUEUE is not None and out.dim() == 2 and out.stride(1) == 1 and dtok.stride(1) == 1 and xb.stride(1) == 1
            and BF16_SHADOW.owns_grad(out)):
        DW_QUEUE.append((torch.cuda.current_stream().cuda_stream, dtok, xb, out))
        return
    K.gemm_bf16(dtok, xb, a_mn=True, b_mn=True, out=out, accumulate=True, split_k=0)


def _colsum_bias(src, out):
    """out (C) += column sums of the bf16 gradient matrix `src` (a bias gradient); queued with the weight gradients when a
    queue is installed and `out` lives in a trainer's flat gradient."""
    if (DW_QUEUE is not None and src.dtype == torch.bfloat16 and src.stride(1) == 1 and src.shape[1] % 8 == 0
            and src.shape[1] <= 2048 and src.stride(0) % 8 == 0 and src.data_ptr() % 16 == 0 and out.is_contiguous()
            and BF16_SHADOW.owns_grad(out)):
        DW_QUEUE.append((torch.cuda.current_stream().cuda_stream, src, None, out))
        return
    K.colsum(src, out)


def flush_dw_queue(final=False):
    """Run the queued weight-gradient products.  Mid-backward (a gradient-bucket boundary) only the products queued on
    the CURRENT stream are run -- their operands are ordered before this point on that stream; `final=True` (after
    backward() has returned and joined its streams) runs everything that is left."""
    q = DW_QUEUE
    if not q:
        return
    cur = torch.cuda.current_stream().cuda_stream
    now = [e for e in q if final or e[0] == cur]
    if not now:
        return
    q[:] = [e for e in q if not (final or e[0] == cur)]
    K.gemm_dw_grouped([(a, b, o) for _s, a, b, o in now if b is not None])
    K.colsum_grouped([(a, o) for _s, a, b, o in now if b is None])


def _param_grads(ctx_params, E, dev):
    """(dW_in, db_in, dWo, dbo) accumulation targets: the parameters' own gradient buffers when they
    exist (trainer flat views), else fresh zero tensors that are returned to autograd."""
    w_in, b_in, w_out, b_out = ctx_params
    slots = [_grad_slot(w_in), _grad_slot(b_in), _grad_slot(w_out), _grad_slot(b_out)]
    shapes = [(3 * E, E), (3 * E,), (E, E), (E,)]
    bufs = [sl if sl is not None else torch.zeros(sh, dtype=torch.float32, device=dev) for sl, sh in zip(slots, shapes)]
    rets = [None if sl is not None else bf_ for sl, bf_ in zip(slots, bufs)]
    return bufs, rets


def _pos_head_grad(d_full, head, L, B, E):
    """Gradient of the learned leading rows of a positional tensor (the rest are constants)."""
    n = head.shape[0]
    g = d_full.view(L, B, E)[:n]
    return g.sum(1, keepdim=True) if head.shape[1] != B else g.clone()


class _MHASelf(torch.autograd.Function):
    """Self-attention block of nn.MultiheadAttention (q = k = x + pos, v = x; head_dim 64):
    in-proj GEMMs write Q/K/V straight into (B, h, L, 64) (head-split epilogue), the whole
    score/softmax/dropout/value chain is ONE fused tcgen05 kernel, O lands token-major (head-merge)
    for the out-proj GEMM.  Backward: the fused attention backward writes dQ|dK|dV side by side
    into ONE (rows, 3E) buffer, so d(x + pos) is a single K = 2E GEMM and dW_in in-place-layout
    GEMMs that accumulate straight into the parameters' gradient buffers; nothing is permuted.
    `pos` is used as a constant; `pos_head` (n, B|1, E), if given, is the learned tensor whose
    values are the first n rows of `pos` and receives their gradient."""

    @staticmethod
    def forward(ctx, x, pos, pos_head, w_in, b_in, w_out, b_out, nh, kpm, p_drop, xqk_hint=None, xv_hint=None):
        L, B, E = x.shape
        dev = x.device
        bf = torch.bfloat16
        xqk_b = xqk_hint if xqk_hint is not None else _tok_bf16(x, pos, L, B, E)
        xv_b = (xv_hint if xv_hint is not None else _tok_bf16(x, None, L, B, E)) if pos is not None else xqk_b
        wb, wo_b = _wb(w_in), _wb(w_out)
        Z = B * nh
        # ONE in-projection launch: output columns [0, 2E) = Q | K from bf16(x + pos), [2E, 3E) = V from bf16(x) (second A
        # operand), written head-split into the three (B, nh, L, 64) parts of one buffer
        qkv = torch.empty((3, Z * L, 64), dtype=bf, device=dev)
        Qh, Kh, Vh = qkv[0], qkv[1], qkv[2]
        K.gemm_ex(L * B, 3 * E, E, 1, xqk_b, False, 0, wb, False, 0, qkv, c_mode=1, hs=(B, nh, L), ldc=64, bias=b_in,
                  a2=xv_b if xv_b is not xqk_b else None, a2_from_col=2 * E, hs_parts=(E, Z * L * 64))
        O_tok, lse, kpm_u8, aux = _attn_core_fwd(Qh, Kh, Vh, L, L, B, nh, kpm, p_drop, dev)
        out = K.gemm_bf16(O_tok, wo_b, bias=b_out)
        ctx.save_for_backward(xqk_b, xv_b, wb, wo_b, Qh, Kh, Vh, lse, kpm_u8, O_tok, pos_head)
        ctx.aux, ctx.dims = aux, (L, B, E, nh, p_drop, pos is not None, None if pos is None else tuple(pos.shape))
        ctx.params = (w_in, b_in, w_out, b_out)
        return out.view(L, B, E)

    @staticmethod
    def backward(ctx, dout):
        xqk_b, xv_b, wb, wo_b, Qh, Kh, Vh, lse, kpm_u8, O_tok, pos_head = ctx.saved_tensors
        L, B, E, nh, p_drop, has_pos, pos_shape = ctx.dims
        dev, bf = dout.device, torch.bfloat16
        Z = B * nh
        (dW_in, db_in, dWo, dbo), rets = _param_grads(ctx.params, E, dev)
        dout_b, bias_done = _grad_bf16(dout, L * B, E, ctx.params[3])
        _dw(dout_b, O_tok, dWo)
        if not bias_done:
            K.colsum(dout_b, dbo)
        dOh = torch.empty((Z * L, 64), dtype=bf, device=dev)
        K.gemm_ex(L * B, E, E, 1, dout_b, False, 0, wo_b, True, 0, dOh, c_mode=1, hs=(B, nh, L), ldc=64)
        buf = torch.empty((L * B, 3 * E), dtype=bf, device=dev)  # [dQ | dK | dV], token-major
        _attn_core_bwd(dOh, Qh, Kh, Vh, O_tok, lse, kpm_u8, L, L, B, nh, ctx.aux, p_drop, buf[:, :E], buf[:, E:2 * E],
                       buf[:, 2 * E:])
        _dw(buf[:, :2 * E], xqk_b, dW_in[:2 * E])
        _dw(buf[:, 2 * E:], xv_b, dW_in[2 * E:])
        _colsum_bias(buf, db_in)
        dpos = dhead = None
        need_pos = has_pos and ctx.needs_input_grad[1]
        need_head = pos_head is not None and ctx.needs_input_grad[2]
        if need_pos:
            d_qk = K.gemm_bf16(buf[:, :2 * E], wb[:2 * E], b_mn=True)          # d(x + pos), K = 2E
            dx = K.gemm_bf16(buf[:, 2 * E:], wb[2 * E:], b_mn=True)
            if need_head:
                dhead = _pos_head_grad(d_qk, pos_head, L, B, E)
            dx += d_qk
            dpos = d_qk.view(L, B, E)
            if pos_shape[1] != B:
                dpos = dpos.sum(1, keepdim=True)
        else:
            dx = K.gemm_bf16(buf, wb, b_mn=True)                                # single K = 3E GEMM
            if need_head:  # d(x + pos) restricted to the learned leading rows: a tiny GEMM
                n = pos_head.shape[0]
                dhead = _pos_head_grad(K.gemm_bf16(buf[: n * B, :2 * E], wb[:2 * E], b_mn=True), pos_head, n, B, E)
        return (dx.view(L, B, E), dpos, dhead, *rets, None, None, None, None, None)


class _MHACross(torch.autograd.Function):
    """Cross-attention block (q = x + qpos, k = mem + mpos, v = mem; head_dim 64), same machinery;
    `mpos_head`: learned leading rows of the constant `mpos` (see _MHASelf)."""

    @staticmethod
    def forward(ctx, x, qpos, mem, mpos, mpos_head, w_in, b_in, w_out, b_out, nh, kpm, p_drop, xq_hint=None, xk_hint=None,
                xv_hint=None):
        L, B, E = x.shape
        S = mem.shape[0]
        dev, bf = x.device, torch.bfloat16
        xq_b = xq_hint if xq_hint is not None else _tok_bf16(x, qpos, L, B, E)
        xk_b = xk_hint if xk_hint is not None else _tok_bf16(mem, mpos, S, B, E)
        xv_b = (xv_hint if xv_hint is not None else _tok_bf16(mem, None, S, B, E)) if mpos is not None else xk_b
        wb, wo_b = _wb(w_in), _wb(w_out)
        Z = B * nh
        Qh = torch.empty((Z * L, 64), dtype=bf, device=dev)
        Kh = torch.empty((Z * S, 64), dtype=bf, device=dev)
        Vh = torch.empty((Z * S, 64), dtype=bf, device=dev)
        K.gemm_ex(L * B, E, E, 1, xq_b, False, 0, wb[:E], False, 0, Qh, c_mode=1, hs=(B, nh, L), ldc=64, bias=b_in[:E])
        K.gemm_ex(S * B, E, E, 1, xk_b, False, 0, wb[E:2 * E], False, 0, Kh, c_mode=1, hs=(B, nh, S), ldc=64,
                  bias=b_in[E:2 * E])
        K.gemm_ex(S * B, E, E, 1, xv_b, False, 0, wb[2 * E:], False, 0, Vh, c_mode=1, hs=(B, nh, S), ldc=64,
                  bias=b_in[2 * E:])
        O_tok, lse, kpm_u8, aux = _attn_core_fwd(Qh, Kh, Vh, L, S, B, nh, kpm, p_drop, dev)
        out = K.gemm_bf16(O_tok, wo_b, bias=b_out)
        ctx.save_for_backward(xq_b, xk_b, xv_b, wb, wo_b, Qh, Kh, Vh, lse, kpm_u8, O_tok, mpos_head)
        ctx.aux = aux
        ctx.dims = (L, S, B, E, nh, p_drop, None if qpos is None else tuple(qpos.shape), None if mpos is None else tuple(mpos.shape))
        ctx.params = (w_in, b_in, w_out, b_out)
        return out.view(L, B, E)

    @staticmethod
    def backward(ctx, dout):
        xq_b, xk_b, xv_b, wb, wo_b, Qh, Kh, Vh, lse, kpm_u8, O_tok, mpos_head = ctx.saved_tensors
        L, S, B, E, nh, p_drop, qpos_shape, mpos_shape = ctx.dims
        dev, bf = dout.device, torch.bfloat16
        Z = B * nh
        (dW_in, db_in, dWo, dbo), rets = _param_grads(ctx.params, E, dev)
        dout_b, bias_done = _grad_bf16(dout, L * B, E, ctx.params[3])
        _dw(dout_b, O_tok, dWo)
        if not bias_done:
            K.colsum(dout_b, dbo)
        dOh = torch.empty((Z * L, 64), dtype=bf, device=dev)
        K.gemm_ex(L * B, E, E, 1, dout_b, False, 0, wo_b, True, 0, dOh, c_mode=1, hs=(B, nh, L), ldc=64)
        dQ_tok = torch.empty((L * B, E), dtype=bf, device=dev)
        kv = torch.empty((S * B, 2 * E), dtype=bf, device=dev)  # [dK | dV]
        _attn_core_bwd(dOh, Qh, Kh, Vh, O_tok, lse, kpm_u8, L, S, B, nh, ctx.aux, p_drop, dQ_tok, kv[:, :E], kv[:, E:])
        _dw(dQ_tok, xq_b, dW_in[:E])
        _dw(kv[:, :E], xk_b, dW_in[E:2 * E])
        _dw(kv[:, E:], xv_b, dW_in[2 * E:])
        K.colsum(dQ_tok, db_in[:E])
        K.colsum(kv, db_in[E:])
        dx = K.gemm_bf16(dQ_tok, wb[:E], b_mn=True)
        dqpos = None
        if qpos_shape is not None and ctx.needs_input_grad[1]:
            dqpos = dx.view(L, B, E) if qpos_shape[1] == B else dx.view(L, B, E).sum(1, keepdim=True)
        dmem = dmpos = dhead = None
        need_mpos = mpos_shape is not None and ctx.needs_input_grad[3]
        need_head = mpos_head is not None and ctx.needs_input_grad[4]
        if need_mpos:
            d_k = K.gemm_bf16(kv[:, :E], wb[E:2 * E], b_mn=True)
            if need_head:
                dhead = _pos_head_grad(d_k, mpos_head, S, B, E)
            if ctx.needs_input_grad[2]:
                dmem = (d_k + K.gemm_bf16(kv[:, E:], wb[2 * E:], b_mn=True)).view(S, B, E)
            dmpos = d_k.view(S, B, E) if mpos_shape[1] == B else d_k.view(S, B, E).sum(1, keepdim=True)
        else:
            if ctx.needs_input_grad[2]:
                dmem = K.gemm_bf16(kv, wb[E:], b_mn=True).view(S, B, E)          # single K = 2E GEMM
            if need_head:  # d(mem + mpos) restricted to the learned leading rows: a tiny GEMM
                n = mpos_head.shape[0]
                d_k = K.gemm_bf16(kv[: n * B, :E], wb[E:2 * E], b_mn=True)
                dhead = _pos_head_grad(d_k, mpos_head, n, B, E)
        return (dx.view(L, B, E), dqpos, dmem, dmpos, dhead, *rets, None, None, None, None, None, None)


# ------------------------------------------------------------------------------------------------
# stage markers: NVTX ranges (PCM_NVTX=1, for ncu --nvtx / Nsight) and optional CUDA-event stage timing
# (tools/stage_times.py).  A no-op otherwise.
# ------------------------------------------------------------------------------------------------
_NVTX = bool(int(os.environ.get("PCM_NVTX", "0")))
STAGE_EVENTS = None  # set to a list by tools/stage_times.py: [(name, start event, end event)]


class stage:
    """`with PF.stage("encoder"):` -- marks one stage of the step on the current stream."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)
        if STAGE_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if STAGE_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STAGE_EVENTS.append((self.name, self.e0, e1))
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


# Gradient-bucket boundaries: the policy marks the activations whose gradient becoming available means "every parameter
# gradient of bucket <tag> is final" (e.g. d(memory): the whole decoder has run its backward); the trainer installs a
# callback that starts that bucket's all-reduce while the rest of the backward is still running.
GRAD_BOUNDARY_CB = None


def grad_boundary(t, tag):
    if torch.is_tensor(t) and t.requires_grad:
        def _hook(_g, tag=tag):
            cb = GRAD_BOUNDARY_CB
            if cb is not None:
                cb(tag)
            return None

        t.register_hook(_hook)
    return t


_PTR_CACHE = {}


def _ptr_table(ptrs, device):
    """Device int64 array of addresses (cached per address tuple: under a trainer the flat buffers never move, so the
    table is built once, before any CUDA-graph capture).  Returns None when it would have to be built during a capture."""
    key = (tuple(ptrs), str(device))
    t = _PTR_CACHE.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        if len(_PTR_CACHE) > 256:
            _PTR_CACHE.clear()
        t = _PTR_CACHE[key] = torch.tensor(list(ptrs), dtype=torch.int64, device=device)
    return t


class _MemKV(torch.autograd.Function):
    """K / V projections of the encoder memory for ALL decoder layers in one launch (the reference recomputes
    `multihead_attn`'s k = W_k (memory + pos), v = W_v memory inside every layer, transformer.py:317-346: 14 GEMMs of
    32 960 x 512 x 512 at cfg-2).  The layers' weight slices are stacked (K blocks, then V blocks) by one gather kernel;
    ONE tcgen05 GEMM with N = 2 n E reads bf16(memory + pos) for the K half and bf16(memory) for the V half and writes
    the 2n head-split (B, h, S, 64) tensors.  Backward: every layer's fused attention backward writes its token-major
    dK / dV into a column block of one shared (S B, 2 n E) buffer; d(memory) is then ONE K = 2 n E GEMM (instead of n
    GEMMs plus n-1 full-size adds), the weight gradients two GEMMs into a stacked buffer that one kernel adds into the
    layers' gradient slots."""

    @staticmethod
    def forward(ctx, mem, mpos, mpos_head, xk_hint, xv_hint, nh, shared, *params):
        from ._lib import check, current_stream, lib, ptr

        S, B, E = mem.shape
        n = len(params) // 2
        dev, bf = mem.device, torch.bfloat16
        xk_b = xk_hint if xk_hint is not None else _tok_bf16(mem, mpos, S, B, E)
        xv_b = (xv_hint if xv_hint is not None else _tok_bf16(mem, None, S, B, E)) if mpos is not None else xk_b
        wbs = [_wb(params[2 * l]) for l in range(n)]
        bs = [params[2 * l + 1] for l in range(n)]
        w_slices = [w[E:2 * E] for w in wbs] + [w[2 * E:] for w in wbs]
        b_slices = [b[E:2 * E] for b in bs] + [b[2 * E:] for b in bs]
        Wkv = torch.empty((2 * n * E, E), dtype=bf, device=dev)
        bkv = torch.empty(2 * n * E, dtype=torch.float32, device=dev)
        tw = _ptr_table([w.data_ptr() for w in w_slices], dev)
        tb = _ptr_table([b.data_ptr() for b in b_slices], dev)
        if tw is not None and tb is not None:
            check(lib.pcm_gather_slices(2 * n, E * E * 2, ptr(tw), ptr(Wkv), current_stream()), "pcm_gather_slices")
            check(lib.pcm_gather_slices(2 * n, E * 4, ptr(tb), ptr(bkv), current_stream()), "pcm_gather_slices")
        else:
            torch.cat(w_slices, 0, out=Wkv)
            torch.cat([b.detach() for b in b_slices], 0, out=bkv)
        Z = B * nh
        KV = torch.empty((2 * n, Z * S, 64), dtype=bf, device=dev)
        K.gemm_ex(S * B, 2 * n * E, E, 1, xk_b, False, 0, Wkv, False, 0, KV, c_mode=1, hs=(B, nh, S), ldc=64, bias=bkv,
                  a2=xv_b if xv_b is not xk_b else None, a2_from_col=n * E, hs_parts=(E, Z * S * 64))
        ctx.save_for_backward(xk_b, xv_b, Wkv, mpos_head)
        ctx.dims = (S, B, E, n, mpos is not None, None if mpos is None else tuple(mpos.shape))
        ctx.params, ctx.shared = params, shared
        return KV

    @staticmethod
    def backward(ctx, dKV):
        from ._lib import check, current_stream, lib, ptr

        xk_b, xv_b, Wkv, mpos_head = ctx.saved_tensors
        S, B, E, n, has_pos, mpos_shape = ctx.dims
        params, shared = ctx.params, ctx.shared
        G = shared.pop("G", None)
        written = shared.pop("written", set())
        none = (None,) * (7 + 2 * n)
        if G is None:
            return none
        dev = G.device
        for l in range(n):  # layers whose backward never ran contribute nothing
            if l not in written:
                G[:, l * E:(l + 1) * E].zero_()
                G[:, (n + l) * E:(n + l + 1) * E].zero_()
        dmem = K.gemm_bf16(G, Wkv, b_mn=True).view(S, B, E) if ctx.needs_input_grad[0] else None  # ONE K = 2nE GEMM
        dmpos = dhead = None
        if has_pos and ctx.needs_input_grad[1]:
            d_k = K.gemm_bf16(G[:, :n * E], Wkv[:n * E], b_mn=True)
            dmpos = d_k.view(S, B, E) if mpos_shape[1] == B else d_k.view(S, B, E).sum(1, keepdim=True)
        if mpos_head is not None and ctx.needs_input_grad[2]:
            h = mpos_head.shape[0]
            d_k = K.gemm_bf16(G[: h * B, :n * E], Wkv[:n * E], b_mn=True)  # learned leading rows only: a tiny GEMM
            dhead = _pos_head_grad(d_k, mpos_head, h, B, E)
        dW = torch.zeros((2 * n * E, E), dtype=torch.float32, device=dev)
        _dw(G[:, :n * E], xk_b, dW[:n * E])
        _dw(G[:, n * E:], xv_b, dW[n * E:])
        db = torch.zeros(2 * n * E, dtype=torch.float32, device=dev)
        step = 1792 if (2 * n * E) % 1792 == 0 else E  # column slices the vectorised column-sum kernel accepts (<= 2048)
        for c0 in range(0, 2 * n * E, step):
            K.colsum(G[:, c0:c0 + step], db[c0:c0 + step])
        w_slots = [_grad_slot(params[2 * l]) for l in range(n)]
        b_slots = [_grad_slot(params[2 * l + 1]) for l in range(n)]
        grads = [None] * (2 * n)
        tw = tb = None
        if all(s_ is not None for s_ in w_slots + b_slots):
            tw = _ptr_table([s_.data_ptr() + E * E * 4 for s_ in w_slots] + [s_.data_ptr() + 2 * E * E * 4 for s_ in w_slots], dev)
            tb = _ptr_table([s_.data_ptr() + E * 4 for s_ in b_slots] + [s_.data_ptr() + 2 * E * 4 for s_ in b_slots], dev)
        if tw is not None and tb is not None:
            check(lib.pcm_add_slices(2 * n, E * E, ptr(tw), ptr(dW), current_stream()), "pcm_add_slices")
            check(lib.pcm_add_slices(2 * n, E, ptr(tb), ptr(db), current_stream()), "pcm_add_slices")
        else:
            for l in range(n):
                gw = torch.zeros((3 * E, E), dtype=torch.float32, device=dev)
                gw[E:2 * E], gw[2 * E:] = dW[l * E:(l + 1) * E], dW[(n + l) * E:(n + l + 1) * E]
                gb = torch.zeros(3 * E, dtype=torch.float32, device=dev)
                gb[E:2 * E], gb[2 * E:] = db[l * E:(l + 1) * E], db[(n + l) * E:(n + l + 1) * E]
                if w_slots[l] is not None:
                    w_slots[l].add_(gw)
                else:
                    grads[2 * l] = gw
                if b_slots[l] is not None:
                    b_slots[l].add_(gb)
                else:
                    grads[2 * l + 1] = gb
        return (dmem, dmpos, dhead, None, None, None, None, *grads)


def memory_kv(mhas, mem, mem_pos, mem_pos_head=None):
    """Project the encoder memory to the keys / values of every decoder layer's cross-attention at once.  `mhas`: the
    layers' nn.MultiheadAttention parameter containers.  Returns an opaque handle for `multi_head_attention(...,
    memkv=(handle, layer_index))`, or None when the shapes are outside the fused kernels (composed path)."""
    S, B, E = mem.shape
    h = mhas[0].num_heads
    if not (mem.is_cuda and E // h == 64 and E % 128 == 0 and all(m.in_proj_weight is not None for m in mhas)):
        return None
    shared = {"consumers": 0}
    xk_hint = _act_bf16(mem, S * B, E) if mem_pos is None else _act_pos_bf16(mem, mem_pos, S * B, E)
    params = []
    for m in mhas:
        params += [m.in_proj_weight, m.in_proj_bias]
    KV = _MemKV.apply(mem, mem_pos, mem_pos_head, xk_hint, _act_bf16(mem, S * B, E), h, shared, *params)
    return KV, shared, len(mhas), S


class _MHACrossKV(torch.autograd.Function):
    """Cross-attention block on keys / values precomputed by `_MemKV` (layer `li` of `n`): Q in-projection, fused
    attention, out-projection.  Backward writes dK / dV token-major into this layer's column blocks of the shared
    (S B, 2 n E) buffer; the K / V weight gradients and d(memory) are formed once for all layers in `_MemKV.backward`."""

    @staticmethod
    def forward(ctx, x, qpos, KV, li, n, S, shared, w_in, b_in, w_out, b_out, nh, kpm, p_drop, xq_hint=None):
        L, B, E = x.shape
        dev, bf = x.device, torch.bfloat16
        xq_b = xq_hint if xq_hint is not None else _tok_bf16(x, qpos, L, B, E)
        wb, wo_b = _wb(w_in), _wb(w_out)
        Z = B * nh
        Qh = torch.empty((Z * L, 64), dtype=bf, device=dev)
        K.gemm_ex(L * B, E, E, 1, xq_b, False, 0, wb[:E], False, 0, Qh, c_mode=1, hs=(B, nh, L), ldc=64, bias=b_in[:E])
        Kh, Vh = KV[li], KV[n + li]
        O_tok, lse, kpm_u8, aux = _attn_core_fwd(Qh, Kh, Vh, L, S, B, nh, kpm, p_drop, dev)
        out = K.gemm_bf16(O_tok, wo_b, bias=b_out)
        ctx.save_for_backward(xq_b, wb, wo_b, Qh, KV, lse, kpm_u8, O_tok)
        ctx.aux = aux
        ctx.dims = (L, S, B, E, nh, p_drop, li, n, None if qpos is None else tuple(qpos.shape))
        ctx.params, ctx.shared = (w_in, b_in, w_out, b_out), shared
        shared["consumers"] += 1
        return out.view(L, B, E)

    @staticmethod
    def backward(ctx, dout):
        xq_b, wb, wo_b, Qh, KV, lse, kpm_u8, O_tok = ctx.saved_tensors
        L, S, B, E, nh, p_drop, li, n, qpos_shape = ctx.dims
        w_in, b_in, w_out, b_out = ctx.params
        shared = ctx.shared
        dev, bf = dout.device, torch.bfloat16
        Z = B * nh
        slots = [_grad_slot(w_in), _grad_slot(b_in), _grad_slot(w_out), _grad_slot(b_out)]
        dW_in = slots[0] if slots[0] is not None else torch.zeros((3 * E, E), dtype=torch.float32, device=dev)
        db_in = slots[1] if slots[1] is not None else torch.zeros(3 * E, dtype=torch.float32, device=dev)
        dWo = slots[2] if slots[2] is not None else torch.zeros((E, E), dtype=torch.float32, device=dev)
        dbo = slots[3] if slots[3] is not None else torch.zeros(E, dtype=torch.float32, device=dev)
        dout_b, bias_done = _grad_bf16(dout, L * B, E, b_out)
        _dw(dout_b, O_tok, dWo)
        if not bias_done:
            K.colsum(dout_b, dbo)
        dOh = torch.empty((Z * L, 64), dtype=bf, device=dev)
        K.gemm_ex(L * B, E, E, 1, dout_b, False, 0, wo_b, True, 0, dOh, c_mode=1, hs=(B, nh, L), ldc=64)
        first = "G" not in shared
        if first:
            shared["G"] = torch.empty((S * B, 2 * n * E), dtype=bf, device=dev)
            shared["written"] = set()
        G = shared["G"]
        dQ_tok = torch.empty((L * B, E), dtype=bf, device=dev)
        _attn_core_bwd(dOh, Qh, KV[li], KV[n + li], O_tok, lse, kpm_u8, L, S, B, nh, ctx.aux, p_drop, dQ_tok,
                       G[:, li * E:(li + 1) * E], G[:, (n + li) * E:(n + li + 1) * E])
        shared["written"].add(li)
        _dw(dQ_tok, xq_b, dW_in[:E])
        _colsum_bias(dQ_tok, db_in[:E])
        dx = K.gemm_bf16(dQ_tok, wb[:E], b_mn=True)
        dqpos = None
        if qpos_shape is not None and ctx.needs_input_grad[1]:
            dqpos = dx.view(L, B, E) if qpos_shape[1] == B else dx.view(L, B, E).sum(1, keepdim=True)
        # the gradient autograd carries for KV is only a token: the first layer to run hands over the shared buffer
        # (reinterpreted in KV's shape), the others return None, so autograd never adds anything up
        dKV = G.view(KV.shape) if first else None
        rets = [None if sl is not None else t for sl, t in zip(slots, (dW_in, db_in, dWo, dbo))]
        return (dx.view(L, B, E), dqpos, dKV, None, None, None, None, *rets, None, None, None, None)


def multi_head_attention(mha, x, pos, mem=None, mem_pos=None, key_padding_mask=None, training=False, pos_head=None,
                         mem_pos_head=None, memkv=None):
    """nn.MultiheadAttention semantics of the reference's call sites (seq-first (L, B, E) tensors;
    only the attended output is returned -- the reference discards the averaged weights it asks
    for, transformer.py:246-248):
        self-attention :  q = k = x + pos,          v = x      (mem is None)
        cross-attention:  q = x + pos, k = mem + mem_pos, v = mem
    `pos_head` / `mem_pos_head` (n, B, E): when the positional tensor is [learned rows ; constant
    rows] (Transformer.forward concatenates additional_pos_embed with the sine embedding), pass the
    detached concatenation as `pos` / `mem_pos` and the learned rows here: only their gradient is
    formed instead of a full (S, B, E) tensor per layer."""
    L, B, E = x.shape
    h = mha.num_heads
    d = E // h
    p = mha.dropout if training else 0.0
    if d == 64 and E % 128 == 0:
        # bf16 operand copies left on the activations by the LayerNorm that produced them
        xq_hint = _act_bf16(x, L * B, E) if pos is None else _act_pos_bf16(x, pos, L * B, E)
        if mem is None:
            out = _MHASelf.apply(x, pos, pos_head, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight,
                                 mha.out_proj.bias, h, key_padding_mask, p, xq_hint, _act_bf16(x, L * B, E))
        elif memkv is not None:  # keys / values of this layer were projected together with all other layers' (memory_kv)
            (KV, shared, n_layers, S), li = memkv
            out = _MHACrossKV.apply(x, pos, KV, li, n_layers, S, shared, mha.in_proj_weight, mha.in_proj_bias,
                                    mha.out_proj.weight, mha.out_proj.bias, h, key_padding_mask, p, xq_hint)
        else:
            S = mem.shape[0]
            xk_hint = _act_bf16(mem, S * B, E) if mem_pos is None else _act_pos_bf16(mem, mem_pos, S * B, E)
            out = _MHACross.apply(x, pos, mem, mem_pos, mem_pos_head, mha.in_proj_weight, mha.in_proj_bias,
                                  mha.out_proj.weight, mha.out_proj.bias, h, key_padding_mask, p, xq_hint, xk_hint,
                                  _act_bf16(mem, S * B, E))
        out._pcm_bias = mha.out_proj.bias  # the LayerNorm that consumes `out` forms this bias gradient (colsum of dx)
        return out
    if pos_head is not None:  # restore the differentiable concatenation for the composed path
        pos = torch.cat([pos_head.expand(-1, B, -1), pos[pos_head.shape[0]:]], dim=0)
    if mem_pos_head is not None:
        mem_pos = torch.cat([mem_pos_head.expand(-1, B, -1), mem_pos[mem_pos_head.shape[0]:]], dim=0)
    # head sizes other than 64 (test fixtures only): composed from the same linear() + library bmm
    query = x if pos is None else x + pos
    if mem is None:
        key, value = query, x
    else:
        key, value = (mem if mem_pos is None else mem + mem_pos), mem
    S = key.shape[0]
    w, b = mha.in_proj_weight, mha.in_proj_bias
    q = linear(query, w[:E], b[:E])
    k = linear(key, w[E: 2 * E], b[E: 2 * E])
    v = linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(L, B, h, d).permute(1, 2, 0, 3)
    k = k.reshape(S, B, h, d).permute(1, 2, 0, 3)
    v = v.reshape(S, B, h, d).permute(1, 2, 0, 3)
    o = attention(q, k, v, key_padding_mask, mha.dropout, training)
    o = o.permute(2, 0, 1, 3).reshape(L, B, E)
    return linear(o, mha.out_proj.weight, mha.out_proj.bias)


class _AddDropoutLN(torch.autograd.Function):
    """y = LayerNorm(res + dropout(x)) in one kernel each way (csrc/layernorm.cu).  Optional
    by-products (non-differentiable): bf16(y) and bf16(y + pos), the operands of the next sub-block's
    GEMMs; the backward leaves bf16(dx) for the sub-block backward that consumes dx."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps, p_drop, want_bf16, pos, x_bias=None, x_exclusive=False):
        """`x_bias`: the bias Parameter of the linear layer that produced `x` (out_proj.bias / linear2.bias): its gradient
        is colsum(dx), accumulated by the backward kernel itself when the parameter owns a flat-gradient slot."""
        shape = res.shape
        C = shape[-1]
        res2 = res.reshape(-1, C)
        x2 = x.reshape(-1, C) if x is not None else None
        seed_base = DROPOUT_RNG.base_for(res.device) if p_drop > 0 else None
        seed = DROPOUT_RNG.next_offset() if p_drop > 0 else 0
        pos2, div = None, 1
        if pos is not None:  # (L, B, C) or (L, 1, C) row-broadcast over the batch of (L, B, C) activations
            div = 1 if pos.shape[1] == shape[1] else shape[1]
            pos2 = pos.reshape(-1, C)
            pos2 = pos2 if pos2.is_contiguous() else pos2.contiguous()
        y, yb, h, mean, rstd, ypb = K.add_dropout_ln_fwd(x2, res2, gamma, beta, eps, p_drop, seed_base, seed,
                                                         want_bf16=want_bf16, pos=pos2, pos_row_div=div)
        ctx.save_for_backward(h, mean, rstd, gamma)
        ctx.cfg = (p_drop, seed_base, seed, x is not None, shape)
        ctx.params = (gamma, beta, x_bias if x is not None else None)
        ctx.x_exclusive = bool(x_exclusive)
        yv = y.view(shape)
        # second handle on the same storage (not an autograd view of the first): consumers that use y as the
        # RESIDUAL operand of the next LayerNorm take this one, so the two gradient contributions of y arrive
        # here separately and are summed inside the backward kernel instead of by an autograd add kernel
        y_res = torch.empty(0, dtype=y.dtype, device=y.device).set_(y.untyped_storage(), y.storage_offset(), shape,
                                                                    yv.stride())
        outs = (yv, yb, ypb, y_res)
        ctx.mark_non_differentiable(*[t for t in outs[1:3] if t is not None])
        ctx.set_materialize_grads(False)  # no zero-filled "gradients" for the bf16 by-products / unused handles
        return outs

    @staticmethod
    def backward(ctx, dy, _dyb=None, _dypb=None, dy_res=None):
        if dy is None:
            dy, dy_res = dy_res, None
        if dy is None:
            return (None,) * 10
        h, mean, rstd, gamma = ctx.saved_tensors
        p_drop, seed_base, seed, has_x, shape = ctx.cfg
        dy2 = dy.reshape(h.shape)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dyr = None
        if dy_res is not None:
            dyr = dy_res.reshape(h.shape)
            if not dyr.is_contiguous():
                dyr = dyr.contiguous()
        g_slot, b_slot = _grad_slot(ctx.params[0]), _grad_slot(ctx.params[1])
        x_bias = ctx.params[2]
        xb_slot = _grad_slot(x_bias) if x_bias is not None else None
        # x tagged with its producer's bias (`_pcm_bias`: the fused attention block / the FFN) and declared by the caller to
        # feed nothing but this LayerNorm: that producer's backward takes bf16(dx) from the side channel and never reads
        # the fp32 tensor, which is then left unwritten (placeholder)
        bf16_only = has_x and x_bias is not None and ctx.x_exclusive and not _NO_LN_DX_BF16_ONLY
        dres, dx, dgamma, dbeta, dxb = K.add_dropout_ln_bwd(dy2, h, mean, rstd, gamma, p_drop, seed_base, seed, has_x,
                                                            dgamma=g_slot, dbeta=b_slot, want_dx_bf16=has_x, dy_b=dyr,
                                                            dx_colsum=xb_slot, dx_fp32=not bf16_only)
        _GRAD_BF16.clear()
        dx_out = None
        if has_x:
            dx_out = dx.view(shape)
            _GRAD_BF16[dx_out.data_ptr()] = (dxb, dx_out, x_bias if xb_slot is not None else None)
            if bf16_only:
                for k in [k for k, r in _DX_UNWRITTEN.items() if r() is None]:
                    del _DX_UNWRITTEN[k]
                _DX_UNWRITTEN[dx_out.data_ptr()] = weakref.ref(dx_out)
        return (dx_out, dres.view(shape), None if g_slot is not None else dgamma, None if b_slot is not None else dbeta,
                None, None, None, None, None, None)


def add_dropout_layernorm(x, residual, norm, p, training, cast=False, cast_pos=None, x_exclusive=False):
    """LayerNorm(residual + dropout(x)) -- the post-LN epilogue of every transformer sub-block.
    `x` may be None (plain LayerNorm of `residual`).  `cast` / `cast_pos`: also produce the bf16
    operand copies bf16(y) / bf16(y + cast_pos) that the NEXT sub-block's GEMMs read (attached to the
    returned tensor; see _act_bf16).  `x_exclusive`: the caller guarantees that `x` (a fused attention block's or FFN's
    output) is used by this call only, so its fp32 gradient need not be materialised (the producer's backward reads the
    bf16 copy)."""
    _need_cuda(residual)
    C = residual.shape[-1]
    p_eff = p if (training and p > 0) else 0.0
    if C % 128 == 0 and C <= 1024 and residual.dtype == torch.float32:
        xr = x.contiguous() if x is not None else None
        if cast_pos is not None and not (cast_pos.dim() == 3 and cast_pos.shape[0] == residual.shape[0] and residual.dim() == 3
                                         and cast_pos.shape[2] == C and cast_pos.dtype == torch.float32):
            cast_pos = None
        res_in = getattr(residual, "_pcm_res", None)  # the producing LayerNorm's residual-branch handle
        if res_in is None or res_in.shape != residual.shape:
            res_in = residual
        y, yb, ypb, y_res = _AddDropoutLN.apply(xr, res_in.contiguous(), norm.weight, norm.bias, norm.eps, p_eff, bool(cast),
                                                None if cast_pos is None else cast_pos.detach(),
                                                getattr(x, "_pcm_bias", None) if x is not None else None, bool(x_exclusive))
        y._pcm_res = y_res
        if yb is not None:
            y._pcm_bf16 = yb
        if ypb is not None:
            y._pcm_bf16_pos = (cast_pos, ypb)
        return y
    # widths that are not a multiple of 128 (test fixtures only): ATen composition
    if x is not None:
        if p_eff > 0:
            x = F.dropout(x, p_eff, True)
        residual = residual + x
    return F.layer_norm(residual, norm.normalized_shape, norm.weight, norm.bias, norm.eps)


_NO_FUSED_FFN = bool(int(os.environ.get("PCM_NO_FUSED_FFN", "0")))  # A/B switch: GEMM -> dropout -> GEMM


class _FFN(torch.autograd.Function):
    """linear2(dropout(relu(linear1(x)))) (transformer.py:243-247,336-340) as ONE autograd node.  dim_feedforward = 32 (the
    reference configuration): ONE kernel each way (csrc/ffn_fused.cu: the hidden tile never leaves registers; y / dx are
    written once).  Other widths: tcgen05 GEMM with bias+ReLU epilogue -> dropout kernel -> GEMM.  Weight / bias gradients
    go to the parameters' gradient buffers through the grouped queue."""

    @staticmethod
    def forward(ctx, x2, w1, b1, w2, b2, p_drop, xb_hint):
        from ._lib import check, current_stream, lib, ptr

        xb = xb_hint if xb_hint is not None else (x2 if x2.dtype == torch.bfloat16 else K.add_cast_bf16(x2))
        w1b, w2b = _wb(w1), _wb(w2)
        rows, E = xb.shape
        Hd = w1.shape[0]
        seed_base, seed = None, 0
        if p_drop > 0:
            seed_base, seed = DROPOUT_RNG.base_for(x2.device), DROPOUT_RNG.next_offset()
        fused = (not _NO_FUSED_FFN and Hd == 32 and E % 64 == 0 and 64 <= E <= 512 and xb.stride(1) == 1 and xb.stride(0) % 8 == 0
                 and w1b.is_contiguous() and w2b.is_contiguous() and w2.shape[0] == E)
        if fused:
            hd = torch.empty((rows, Hd), dtype=torch.bfloat16, device=xb.device)
            y = torch.empty((rows, E), dtype=torch.float32, device=xb.device)
            check(lib.pcm_ffn32_fwd(rows, E, Hd, ptr(xb), xb.stride(0), ptr(w1b), ptr(b1), ptr(w2b), ptr(b2), float(p_drop),
                                    ptr(seed_base), int(seed), ptr(hd), ptr(y), current_stream()), "pcm_ffn32_fwd")
            h = hd
        else:
            h = K.gemm_bf16(xb, w1b, bias=b1, relu=True, out_dtype=torch.bfloat16)  # (rows, Hd) bf16
            hd = h
            if p_drop > 0:
                hd = torch.empty_like(h)
                check(lib.pcm_ffn_dropout_fwd(h.shape[0], h.shape[1], ptr(h), float(p_drop), ptr(seed_base), int(seed), ptr(hd),
                                              current_stream()), "pcm_ffn_dropout_fwd")
            y = K.gemm_bf16(hd, w2b, bias=b2)
        ctx.save_for_backward(xb, h, hd, w1b, w2b, seed_base)
        ctx.cfg = (float(p_drop), int(seed), fused)
        ctx.params = (w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        from ._lib import check, current_stream, lib, ptr

        xb, h, hd, w1b, w2b, seed_base = ctx.saved_tensors
        p_drop, seed, fused = ctx.cfg
        w1, b1, w2, b2 = ctx.params
        rows, Hd = h.shape
        dev = dy.device
        dyb, b2_done = _grad_bf16(dy, rows, dy.shape[1], b2)
        slot2 = _grad_slot(w2)
        dw2 = slot2 if slot2 is not None else torch.zeros(w2.shape, dtype=torch.float32, device=dev)
        _dw(dyb, hd, dw2)
        slotb2 = _grad_slot(b2)
        db2 = slotb2 if b2_done else K.colsum(dyb, slotb2)
        slotb1 = _grad_slot(b1)
        db1 = slotb1 if slotb1 is not None else torch.zeros(Hd, dtype=torch.float32, device=dev)
        dhb = torch.empty_like(h)
        dx = None
        if fused:
            dx = torch.empty((rows, dy.shape[1]), dtype=torch.float32, device=dev)
            check(lib.pcm_ffn32_bwd(rows, dy.shape[1], Hd, ptr(dyb), dyb.stride(0), ptr(hd), ptr(w1b), ptr(w2b), float(p_drop),
                                    ptr(dhb), ptr(dx), current_stream()), "pcm_ffn32_bwd")
            _colsum_bias(dhb, db1)
        else:
            dhd = K.gemm_bf16(dyb, w2b, b_mn=True)  # (rows, E) x W2(E, Hd) -> (rows, Hd) fp32
            fuse_b1 = Hd <= 256 and 256 % (Hd // 8) == 0  # the gate kernel can form colsum(dh) = db1 itself
            check(lib.pcm_ffn_relu_dropout_bwd_ex(rows, Hd, ptr(dhd), ptr(h), float(p_drop), ptr(seed_base), int(seed), ptr(dhb),
                                                  ptr(db1 if fuse_b1 else None), current_stream()), "pcm_ffn_relu_dropout_bwd_ex")
            if not fuse_b1:
                K.colsum(dhb, db1)
            if ctx.needs_input_grad[0]:
                dx = K.gemm_bf16(dhb, w1b, b_mn=True)
        slot1 = _grad_slot(w1)
        dw1 = slot1 if slot1 is not None else torch.zeros(w1.shape, dtype=torch.float32, device=dev)
        _dw(dhb, xb, dw1)
        return (dx, None if slot1 is not None else dw1, None if slotb1 is not None else db1, None if slot2 is not None else dw2,
                None if slotb2 is not None else db2, None, None)


def feed_forward(x, linear1, linear2, p_drop, training):
    """FFN sub-block on token-major activations x (..., E); `linear1` / `linear2` are the nn.Linear parameter
    containers.  Falls back to two `linear` calls when the shapes are not TMA-legal (never on the reference configs)."""
    _need_cuda(x)
    E, Hd = linear1.weight.shape[1], linear1.weight.shape[0]
    p = float(p_drop) if training else 0.0
    if E % 8 or Hd % 8 or E < 16 or Hd < 16 or linear1.bias is None or linear2.bias is None:
        h = linear(x, linear1.weight, linear1.bias, relu=True)
        return linear(dropout(h, p_drop, training), linear2.weight, linear2.bias)
    x2 = x.reshape(-1, E)
    if x2.stride(-1) != 1 or (x2.stride(0) % 8) or (x2.data_ptr() % 16):
        x2 = x2.contiguous()
    y = _FFN.apply(x2, linear1.weight, linear1.bias, linear2.weight, linear2.bias, p, _act_bf16(x, x2.shape[0], E))
    y = y.view(*x.shape[:-1], linear2.weight.shape[0])
    y._pcm_bias = linear2.bias  # see multi_head_attention
    return y


def dropout(x, p, training):
    return F.dropout(x, p, True) if (training and p > 0) else x


class _SyncBN:
    """SyncBatchNorm switch (reference DDP preset: configs/trainer/ddp.yaml:9 `sync_batchnorm: true`).  When enabled, every
    training-mode BatchNorm of the path (PointNet / projector layers: `_BatchNormReLU`; the set-abstraction head:
    `_SetAbstraction`) all-reduces its (sum, sum of squares, row count) between the statistics kernel and the finalize
    kernel, and its two gradient sums between the backward's reduce and apply kernels -- one small collective per layer
    each way, sequentially dependent, exactly like torch.nn.SyncBatchNorm.  Off (default) = per-rank statistics = the
    reference with `trainer.sync_batchnorm=false`."""

    def __init__(self):
        self.enabled, self.group = False, None

    def active(self):
        import torch.distributed as dist

        return self.enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1


SYNC_BN = _SyncBN()


def set_sync_batchnorm(enabled: bool, process_group=None):
    SYNC_BN.enabled, SYNC_BN.group = bool(enabled), process_group


def _allreduce_stats(rows, count):
    """All-reduce the fp64 statistics rows (k, C) together with the local row count; returns the global count as a (1,)
    fp64 DEVICE tensor (read by the kernels through their n_rows_dev argument: no host synchronisation)."""
    import torch.distributed as dist

    buf = torch.cat([rows.reshape(-1), torch.full((1,), float(count), dtype=torch.float64, device=rows.device)])
    dist.all_reduce(buf, group=SYNC_BN.group)
    rows.copy_(buf[:-1].view_as(rows))
    return buf[-1:]


class _BatchNormReLU(torch.autograd.Function):
    """ReLU(BatchNorm1d(y)) over rows (csrc/batchnorm.cu): statistics pass + apply pass forward, reduce pass +
    apply pass backward; emits the bf16 operand of the next GEMM (forward) and bf16(dy) for the producing
    linear's backward GEMMs, accumulates dgamma / dbeta into the parameters' gradient buffers."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, eps, momentum, training, relu, extra=None):
        """`extra`: [(running_mean, running_var, momentum), ...] of further BatchNorm layers that are fed the SAME input in
        training mode and only need their running statistics updated (PDBatchNorm's non-selected conditions)."""
        from ._lib import check, current_stream, lib, ptr

        R, C = y.shape
        st = current_stream()
        stats = _zeros((2, C), torch.float64, y.device)
        n_dev = None
        if training:
            check(lib.pcm_bn_stats(R, C, ptr(y), ptr(stats), st), "pcm_bn_stats")
            if SYNC_BN.active():
                n_dev = _allreduce_stats(stats, R)
        coef = torch.empty((4, C), dtype=torch.float32, device=y.device)
        if training and extra:
            for rm, rv, mom in extra:
                check(lib.pcm_sa_bn_finalize_ex(C, ptr(stats), float(R), ptr(n_dev), ptr(gamma), ptr(beta), float(eps), float(mom), 1,
                                                ptr(rm), ptr(rv), ptr(coef), st), "pcm_sa_bn_finalize_ex")
        check(lib.pcm_sa_bn_finalize_ex(C, ptr(stats), float(R), ptr(n_dev), ptr(gamma), ptr(beta), float(eps), float(momentum),
                                        int(training), ptr(running_mean), ptr(running_var), ptr(coef), st), "pcm_sa_bn_finalize_ex")
        ctx.n_dev = n_dev
        out = torch.empty_like(y)
        outb = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device)
        check(lib.pcm_bn_apply_relu(R, C, ptr(y), ptr(coef), int(relu), ptr(out), ptr(outb), st), "pcm_bn_apply_relu")
        ctx.save_for_backward(y, coef)
        ctx.cfg = (bool(training), bool(relu))
        ctx.params = (gamma, beta)
        ctx.mark_non_differentiable(outb)
        ctx.set_materialize_grads(False)
        return out, outb

    @staticmethod
    def backward(ctx, dout, _doutb=None):
        from ._lib import check, current_stream, lib, ptr

        if dout is None:
            return (None,) * 10
        y, coef = ctx.saved_tensors
        training, relu = ctx.cfg
        gamma, beta = ctx.params
        R, C = y.shape
        dout = dout.contiguous()
        g_slot, b_slot = _grad_slot(gamma), _grad_slot(beta)
        dg = g_slot if g_slot is not None else torch.zeros(C, dtype=torch.float32, device=y.device)
        db = b_slot if b_slot is not None else torch.zeros(C, dtype=torch.float32, device=y.device)
        gstats = _zeros((2, C), torch.float64, y.device)
        need_dy = ctx.needs_input_grad[0]
        dy = torch.empty_like(y) if need_dy else None
        dyb = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device)
        if ctx.n_dev is not None:  # SyncBatchNorm: global sums for dx, LOCAL sums for the affine gradients
            import torch.distributed as dist

            check(lib.pcm_bn_relu_bwd_reduce(R, C, ptr(dout), ptr(y), ptr(coef), int(relu), ptr(gstats), current_stream()),
                  "pcm_bn_relu_bwd_reduce")
            local = gstats.clone()
            dist.all_reduce(gstats, group=SYNC_BN.group)
            check(lib.pcm_bn_relu_bwd_apply(R, C, ptr(dout), ptr(y), ptr(coef), int(relu), int(training), ptr(gstats), ptr(ctx.n_dev),
                                            ptr(dy), ptr(dyb), None, None, current_stream()), "pcm_bn_relu_bwd_apply")
            db += local[0].float()
            dg += local[1].float()
        else:
            check(lib.pcm_bn_relu_bwd(R, C, ptr(dout), ptr(y), ptr(coef), int(relu), int(training), ptr(gstats), ptr(dy), ptr(dyb),
                                      ptr(dg), ptr(db), current_stream()), "pcm_bn_relu_bwd")
        if dy is not None:
            _GRAD_BF16[dy.data_ptr()] = (dyb, dy)  # consumed (popped) by the producing linear's backward
        return (dy, None if g_slot is not None else dg, None if b_slot is not None else db, None, None, None, None, None, None, None)


_NO_FUSED_BN = bool(int(os.environ.get("PCM_NO_FUSED_BN", "0")))  # A/B switch (ATen composition)


def batchnorm_relu(x, bn, relu=True):
    """ReLU(BatchNorm1d(x)) over rows of x (R, C); training mode uses batch statistics and updates
    the running buffers exactly like nn.BatchNorm1d (momentum, unbiased running_var).  The returned activation
    carries its bf16 copy (`_pcm_bf16`), the operand of the next layer's GEMM."""
    C = x.shape[-1]
    training = bn.training or bn.running_mean is None
    if (not _NO_FUSED_BN and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and C % 4 == 0 and C <= 1024 and bn.affine
            and (bn.momentum is not None or not training)):
        out, outb = _BatchNormReLU.apply(x.contiguous(), bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                                         bn.momentum if bn.momentum is not None else 0.0, training, relu)
        out._pcm_bf16 = outb
        return out
    y = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training,
                     bn.momentum if bn.momentum is not None else 0.0, bn.eps)
    return F.relu(y) if relu else y


class _SetAbstraction(torch.autograd.Function):
    """Fused gather + Linear(3+C->H) + BatchNorm1d + ReLU + max over k (csrc/sa_fused.cu)."""

    @staticmethod
    def forward(ctx, feat, weight, gamma, beta, p, new_p, knn_idx, running_mean, running_var, eps, momentum,
                training, feat_b=None, tok=None, clouds=None):
        """`tok` = (batch, head_rows, pos or None, out_bf16, out_pos_bf16 or None): write the result straight into the
        transformer's seq-first token tensor (S, B, H) -- query b * M + mi -> row (head_rows + mi) * B + b; the head rows
        are left for `fill_head_rows` -- and emit bf16(out) / bf16(out + pos), the operands of the first encoder layer.
        `clouds` = (offset int32 (b), new_offset int32 (b), n_max): per-cloud extents + host-known size bound -> the
        cloud-slice gather kernel (Pf slice of a cloud in shared memory; csrc/sa_fused.cu), generic kernel otherwise."""
        from ._lib import check, current_stream, lib, ptr

        n, C = feat.shape
        H = weight.shape[0]
        m, k = knn_idx.shape
        dev = feat.device
        st = current_stream()
        weight = weight.contiguous()
        featb = feat_b if feat_b is not None else (feat if feat.dtype == torch.bfloat16 else feat.to(torch.bfloat16))
        wfb = weight[:, 3:].to(torch.bfloat16).contiguous()
        Pf = K.gemm_bf16(featb, wfb)  # (n, H) fp32: the only tensor-core work of the layer
        ymax = torch.empty((m, H), dtype=torch.float32, device=dev)
        jmax = torch.empty((m, H), dtype=torch.uint8, device=dev)
        ymin = jmin = None
        stats = _zeros((5, H), torch.float64, dev)
        # gather pass, fastest applicable form first.  Single-extreme forms: sign(BatchNorm scale) = sign(gamma) is known
        # before the statistics, so only the extreme that will be selected is tracked (bit-identical results).
        rc = PCM_EUNSUPPORTED
        if not _NO_SA_SEL:
            if clouds is not None and not _NO_SA_CLOUDS:
                off, noff, n_max = clouds
                rc = lib.pcm_sa_gather_sel_clouds(off.numel(), int(n_max), m, k, H, ptr(Pf), ptr(p), ptr(new_p), ptr(knn_idx),
                                                  ptr(off), ptr(noff), ptr(weight), weight.stride(0), ptr(gamma), ptr(ymax),
                                                  ptr(jmax), ptr(stats), st)
            if rc == PCM_EUNSUPPORTED:
                rc = lib.pcm_sa_gather_sel(m, k, H, ptr(Pf), ptr(p), ptr(new_p), ptr(knn_idx), ptr(weight), weight.stride(0),
                                           ptr(gamma), ptr(ymax), ptr(jmax), ptr(stats), st)
            if rc != PCM_EUNSUPPORTED:
                check(rc, "pcm_sa_gather_sel")
        if rc == PCM_EUNSUPPORTED:
            ymin, jmin = torch.empty_like(ymax), torch.empty_like(jmax)
            if clouds is not None and not _NO_SA_CLOUDS:
                off, noff, n_max = clouds
                rc = lib.pcm_sa_gather_stats_clouds(off.numel(), int(n_max), m, k, H, ptr(Pf), ptr(p), ptr(new_p), ptr(knn_idx),
                                                    ptr(off), ptr(noff), ptr(weight), weight.stride(0), ptr(ymax), ptr(ymin),
                                                    ptr(jmax), ptr(jmin), ptr(stats), st)
                if rc != PCM_EUNSUPPORTED:
                    check(rc, "pcm_sa_gather_stats_clouds")
            if rc == PCM_EUNSUPPORTED:
                check(lib.pcm_sa_gather_stats(m, k, H, ptr(Pf), ptr(p), ptr(new_p), ptr(knn_idx), ptr(weight), weight.stride(0),
                                              ptr(ymax), ptr(ymin), ptr(jmax), ptr(jmin), ptr(stats), st), "pcm_sa_gather_stats")
        coef = torch.empty((4, H), dtype=torch.float32, device=dev)
        n_dev = None
        if training and SYNC_BN.active():  # rows 0-1 (sum y, sum y^2) global; rows 2-4 (sum y dxyz) stay local (backward dW)
            n_dev = _allreduce_stats(stats[:2], m * k)
        ctx.n_dev = n_dev
        check(lib.pcm_sa_bn_finalize_ex(H, ptr(stats), float(m * k), ptr(n_dev), ptr(gamma), ptr(beta), float(eps), float(momentum),
                                        int(training), ptr(running_mean), ptr(running_var), ptr(coef), st),
              "pcm_sa_bn_finalize_ex")
        ctx.tok = None
        if tok is None:
            out = torch.empty((m, H), dtype=torch.float32, device=dev)
            check(lib.pcm_sa_output(m, H, ptr(ymax), ptr(ymin), ptr(jmax), ptr(jmin), ptr(coef), ptr(out), ptr(jmax), st),
                  "pcm_sa_output")  # jsel overwrites jmax in place
            saved_out = out
        else:
            B, head, pos, out_b, out_pb = tok
            per = m // B
            out = torch.empty((head + per, B, H), dtype=torch.float32, device=dev)
            check(lib.pcm_sa_output_tokens(m, H, per, B, head, ptr(ymax), ptr(ymin), ptr(jmax), ptr(jmin), ptr(coef), ptr(pos),
                                           ptr(out), ptr(out_b), ptr(out_pb), ptr(jmax), st), "pcm_sa_output_tokens")
            ctx.tok = (B, head, per)
            # a second handle on the storage for backward: `fill_head_rows` writes the head rows of `out` in place later
            # (which bumps the version counter of `out` itself)
            saved_out = torch.empty(0, dtype=out.dtype, device=dev).set_(out.untyped_storage(), 0, out.shape, out.stride())
        ctx.save_for_backward(featb, wfb, Pf, p, new_p, knn_idx, weight, saved_out, jmax, coef, stats)
        ctx.training = training
        return out

    @staticmethod
    def backward(ctx, dout):
        from ._lib import check, current_stream, lib, ptr

        featb, wfb, Pf, p, new_p, knn_idx, weight, out, jsel, coef, stats = ctx.saved_tensors
        n, H = Pf.shape
        C = featb.shape[1]
        m, k = knn_idx.shape
        dev = Pf.device
        st = current_stream()
        dout = dout.contiguous().float()
        gstats = _zeros((5, H), torch.float64, dev)
        dPf = torch.zeros((n, H), dtype=torch.float32, device=dev)
        if ctx.tok is None:
            check(lib.pcm_sa_bwd_scatter(m, k, H, ptr(dout), ptr(out), ptr(jsel), ptr(knn_idx), ptr(p), ptr(new_p), ptr(coef),
                                         ptr(dPf), ptr(gstats), st), "pcm_sa_bwd_scatter")
        else:
            B, head, per = ctx.tok
            e = _GRAD_EXTRA.pop(dout.data_ptr(), None)  # second gradient of the token tensor (residual branch)
            dout2 = e[0] if (e is not None and e[1].data_ptr() == dout.data_ptr() and e[0].shape == dout.shape) else None
            check(lib.pcm_sa_bwd_scatter_tokens(m, k, H, per, B, head, ptr(dout), ptr(dout2), ptr(out), ptr(jsel), ptr(knn_idx),
                                                ptr(p), ptr(new_p), ptr(coef), ptr(dPf), ptr(gstats), st),
                  "pcm_sa_bwd_scatter_tokens")
        cnt = _zeros((n,), torch.float32, dev)
        sq = _zeros((n, 3), torch.float32, dev)
        sdtot = _zeros((3,), torch.float64, dev)
        check(lib.pcm_sa_edge_stats(m, k, ptr(knn_idx), ptr(p), ptr(new_p), ptr(cnt), ptr(sq), ptr(sdtot), st),
              "pcm_sa_edge_stats")
        dW = _zeros((H, 3 + C), torch.float32, dev)
        ab = torch.empty((2, H), dtype=torch.float32, device=dev)
        dgamma = torch.empty(H, dtype=torch.float32, device=dev)
        dbeta = torch.empty(H, dtype=torch.float32, device=dev)
        local = None
        if ctx.n_dev is not None:  # SyncBatchNorm: alpha / beta' from the global sums, dgamma / dbeta from the local ones
            import torch.distributed as dist

            local = gstats[:2].clone()
            dist.all_reduce(gstats[:2], group=SYNC_BN.group)
        check(lib.pcm_sa_bwd_coef_ex(H, ptr(gstats), ptr(stats), ptr(sdtot), ptr(coef), float(m * k), ptr(ctx.n_dev),
                                     int(ctx.training), ptr(ab), ptr(dW), dW.stride(0), ptr(dgamma), ptr(dbeta), st),
              "pcm_sa_bwd_coef_ex")
        if local is not None:
            dbeta.copy_(local[0])
            dgamma.copy_(local[1])
        dPfb = torch.empty((n, H), dtype=torch.bfloat16, device=dev)
        check(lib.pcm_sa_bwd_dense(n, H, ptr(Pf), ptr(p), ptr(cnt), ptr(sq), ptr(weight), weight.stride(0), ptr(ab),
                                   ptr(dPf), ptr(dPfb), st), "pcm_sa_bwd_dense")
        dfeat = None
        if ctx.needs_input_grad[0]:
            dfeat = K.gemm_bf16(dPfb, wfb, b_mn=True)  # (n, H) x Wf(H, C)
        K.gemm_bf16(dPfb, featb, a_mn=True, b_mn=True, out=dW[:, 3:], accumulate=True, split_k=0)
        return dfeat, dW, dgamma, dbeta, None, None, None, None, None, None, None, None, None, None, None


_NO_SA_CLOUDS = bool(int(os.environ.get("PCM_NO_SA_CLOUDS", "0")))  # A/B switch: gather without the shared-memory cloud slice
_NO_SA_SEL = bool(int(os.environ.get("PCM_NO_SA_SEL", "0")))  # A/B switch: track both extremes (max and min) per channel
PCM_EUNSUPPORTED = -2


def set_abstraction(p, feat, offset, new_p, new_offset, knn_idx, linear_weight, bn, tokens=None, n_max=None):
    """Grouped Linear(3+C -> H, no bias) + BatchNorm1d + ReLU + max over the k neighbours
    (act.py:446-460) as ONE fused operator.  feat (n, C), knn_idx (m, k) int32 (-1 = padding) -> (m, H).
    `tokens` = (batch, head_rows, pos (S, B, H) or None): return the transformer's seq-first token tensor (S, B, H)
    instead, with the point rows filled (see `fill_head_rows` for the rest) and the bf16 operand copies attached.
    `n_max`: host-known upper bound of the cloud sizes (the batch's `n_max` hint) -> cloud-slice kernels."""
    _need_cuda(feat)
    training = bn.training or (bn.running_mean is None)
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = bn.momentum if bn.momentum is not None else 0.0
    H, C = linear_weight.shape[0], feat.shape[1]
    if C % 8 or H % 8:
        # narrow head (pre_sample: Linear(3+6 -> 6), act.py:371-375): zero-pad channels to the TMA / float4
        # granularity.  Padded outputs are identically 0 (y = 0 on every edge -> BN(0) with beta 0 -> 0) and
        # receive zero gradient; autograd slices the real gradients back out of the padded tensors.
        Cp, Hp = -(-C // 8) * 8, -(-H // 8) * 8
        w = F.pad(linear_weight, (0, Cp - C, 0, Hp - H))
        rm = F.pad(bn.running_mean, (0, Hp - H)) if bn.running_mean is not None else None
        rv = F.pad(bn.running_var, (0, Hp - H), value=1.0) if bn.running_var is not None else None
        out = _SetAbstraction.apply(F.pad(feat, (0, Cp - C)), w, F.pad(bn.weight, (0, Hp - H), value=1.0),
                                    F.pad(bn.bias, (0, Hp - H)), p.contiguous(), new_p.contiguous(),
                                    knn_idx.contiguous(), rm, rv, bn.eps, momentum, training)
        if training and rm is not None:
            bn.running_mean.copy_(rm[:H])
            bn.running_var.copy_(rv[:H])
        return out[:, :H]
    tok = None
    if tokens is not None:
        B, head, pos = tokens
        m = knn_idx.shape[0]
        if H % 4 or m % B:
            raise PcmError("token-layout set abstraction needs H % 4 == 0 and the same number of queries per cloud")
        S = head + m // B
        if pos is not None and (tuple(pos.shape) != (S, B, H) or not pos.is_contiguous() or pos.dtype != torch.float32):
            raise PcmError("pos must be a contiguous fp32 (S, B, H) tensor")
        out_b = torch.empty((S, B, H), dtype=torch.bfloat16, device=feat.device)
        out_pb = torch.empty_like(out_b) if pos is not None else None
        tok = (B, head, pos, out_b, out_pb)
    clouds = None
    if (n_max is not None and int(n_max) > 0 and torch.is_tensor(offset) and torch.is_tensor(new_offset)
            and offset.dtype == torch.int32 and new_offset.dtype == torch.int32 and offset.numel() == new_offset.numel()):
        clouds = (offset.contiguous(), new_offset.contiguous(), int(n_max))
    out = _SetAbstraction.apply(feat, linear_weight, bn.weight, bn.bias, p.contiguous(), new_p.contiguous(),
                                knn_idx.contiguous(), bn.running_mean, bn.running_var, bn.eps, momentum, training,
                                _act_bf16(feat, feat.shape[0], C), tok, clouds)
    if tok is not None:
        out._pcm_bf16 = tok[3]
        if tok[4] is not None:
            out._pcm_bf16_pos = (pos, tok[4])
    return out


class _FillHeadRows(torch.autograd.Function):
    """Rows 0 .. head-1 of the seq-first token tensor (S, B, E) = [latent ; proprio / goal] (transformer.py:89-92), written
    IN PLACE by one small kernel (the point rows come from the token-layout set abstraction), together with the bf16
    operand copies of those rows.  Returns two handles on the same storage (see _AddDropoutLN): the second one is what
    the first LayerNorm takes as its residual operand, so the two gradient contributions of the tokens arrive
    separately and are summed inside the set-abstraction backward kernel instead of by an autograd add pass."""

    @staticmethod
    def forward(ctx, tokens, latent, proprio, pos, tok_b, tok_pb):
        from ._lib import check, current_stream, lib, ptr

        S, B, E = tokens.shape
        head = 1 + (proprio.shape[0] if proprio is not None else 0)
        latent2 = latent.reshape(B, E).contiguous()
        prop2 = proprio.reshape(head - 1, B, E).contiguous() if proprio is not None else None
        check(lib.pcm_fill_head_rows(B, head, E, ptr(latent2), ptr(prop2), ptr(pos), ptr(tokens), ptr(tok_b),
                                     ptr(tok_pb if pos is not None else None), current_stream()), "pcm_fill_head_rows")
        ctx.mark_dirty(tokens)
        ctx.head = head
        ctx.shapes = (latent.shape, None if proprio is None else proprio.shape)
        res = torch.empty(0, dtype=tokens.dtype, device=tokens.device).set_(tokens.untyped_storage(), tokens.storage_offset(),
                                                                            tokens.shape, tokens.stride())
        ctx.set_materialize_grads(False)
        return tokens, res

    @staticmethod
    def backward(ctx, d_tok, d_res):
        if d_tok is None:
            d_tok, d_res = d_res, None
        if d_tok is None:
            return (None,) * 6
        head = ctx.head
        lat_shape, prop_shape = ctx.shapes
        d_tok = d_tok.contiguous()
        dh = d_tok[:head]
        if d_res is not None:
            d_res = d_res.contiguous()
            dh = dh + d_res[:head]
            _GRAD_EXTRA[d_tok.data_ptr()] = (d_res, d_tok)  # popped by the set-abstraction backward
        d_lat = dh[0].reshape(lat_shape) if ctx.needs_input_grad[1] else None
        d_prop = dh[1:].reshape(prop_shape) if (prop_shape is not None and ctx.needs_input_grad[2]) else None
        return d_tok, d_lat, d_prop, None, None, None


def fill_head_rows(tokens, latent, proprio, pos=None):
    """Complete the token tensor produced by `set_abstraction(..., tokens=...)`; returns it with the bf16 operand copies
    (`_pcm_bf16`, `_pcm_bf16_pos`) and the residual-branch handle (`_pcm_res`) attached."""
    tok_b = getattr(tokens, "_pcm_bf16", None)
    pe = getattr(tokens, "_pcm_bf16_pos", None)
    tok_pb = pe[1] if (pe is not None and pe[0] is pos) else None
    out, res = _FillHeadRows.apply(tokens, latent, proprio, pos if tok_pb is not None else None, tok_b, tok_pb)
    out._pcm_res = res
    if tok_b is not None:
        out._pcm_bf16 = tok_b
    if tok_pb is not None:
        out._pcm_bf16_pos = (pos, tok_pb)
    return out


class _ActHeadsLoss(torch.autograd.Function):
    """action_head + is_pad_head + masked MSE + KL (act.py:255-291, RLBench :770-825) as one kernel each way
    (csrc/tokens.cu).  Outputs (a_hat, is_pad_hat, losses[3]); `actions is None`: heads only."""

    @staticmethod
    def forward(ctx, hs, Wa, ba, Wp, bp, actions, is_pad, mu, logvar, cfg):
        from ._lib import check, current_stream, lib, ptr

        sig_start, n_pos, w_pos, kl_weight = cfg
        B, Q, E = hs.shape
        A = Wa.shape[0]
        dev = hs.device
        assert hs.stride(2) == 1
        a_hat = torch.empty((B, Q, A), dtype=torch.float32, device=dev)
        pad_hat = torch.empty((B, Q, 1), dtype=torch.float32, device=dev)
        losses = _zeros((3,), torch.float32, dev) if actions is not None else None
        ws = _workspace("heads_loss_ws", (4,), torch.float64, dev, zero=True)  # acc (1 double) + ticket; self-cleaning
        pad_u8 = is_pad.contiguous().view(torch.uint8) if is_pad is not None else None
        act_c = actions.contiguous().float() if actions is not None else None
        mu_c = mu.contiguous() if mu is not None else None
        lv_c = logvar.contiguous() if logvar is not None else None
        L = mu.shape[1] if mu is not None else 0
        check(lib.pcm_act_heads_loss_fwd(B, Q, E, A, L, int(sig_start), int(n_pos), float(w_pos), float(kl_weight), ptr(hs),
                                         hs.stride(0), hs.stride(1), ptr(Wa), ptr(ba), ptr(Wp), ptr(bp), ptr(act_c), ptr(pad_u8),
                                         ptr(mu_c), ptr(lv_c), ptr(a_hat), ptr(pad_hat), ptr(losses), ptr(ws), ptr(ws[1:]),
                                         current_stream()), "pcm_act_heads_loss_fwd")
        ctx.save_for_backward(hs, Wa, Wp, act_c, pad_u8, mu_c, lv_c, a_hat)
        ctx.cfg = cfg
        ctx.params = (Wa, ba, Wp, bp)
        ctx.set_materialize_grads(False)
        if losses is None:
            return a_hat, pad_hat, None, None, None
        return a_hat, pad_hat, losses[0], losses[1], losses[2]

    @staticmethod
    def backward(ctx, g_a_hat, g_pad, g_loss, g_action, g_kl):
        from ._lib import check, current_stream, lib, ptr

        hs, Wa, Wp, act_c, pad_u8, mu_c, lv_c, a_hat = ctx.saved_tensors
        sig_start, n_pos, w_pos, kl_weight = ctx.cfg
        B, Q, E = hs.shape
        A = Wa.shape[0]
        dev = hs.device
        pWa, pba, pWp, pbp = ctx.params
        g_loss, g_action, g_kl = (None if g is None else g.reshape(1).float() for g in (g_loss, g_action, g_kl))
        slots = [_grad_slot(pWa), _grad_slot(pba)]
        dWa = slots[0] if slots[0] is not None else torch.zeros_like(Wa)
        dba = slots[1] if slots[1] is not None else torch.zeros(A, dtype=torch.float32, device=dev)
        dWp = dbp = None
        pslots = [None, None]
        if g_pad is not None:
            pslots = [_grad_slot(pWp), _grad_slot(pbp)]
            dWp = pslots[0] if pslots[0] is not None else torch.zeros_like(Wp)
            dbp = pslots[1] if pslots[1] is not None else torch.zeros(1, dtype=torch.float32, device=dev)
            g_pad = g_pad.contiguous().float()
        if g_a_hat is not None:
            g_a_hat = g_a_hat.contiguous().float()
        d_hs = torch.empty((Q, B, E), dtype=torch.float32, device=dev)  # the decoder's (Q, B, E) memory order
        dmu = torch.empty_like(mu_c) if mu_c is not None else None
        dlv = torch.empty_like(lv_c) if lv_c is not None else None
        L = mu_c.shape[1] if mu_c is not None else 0
        check(lib.pcm_act_heads_loss_bwd(B, Q, E, A, L, int(sig_start), int(n_pos), float(w_pos), float(kl_weight), ptr(hs),
                                         hs.stride(0), hs.stride(1), ptr(Wa), ptr(Wp), ptr(act_c), ptr(pad_u8), ptr(mu_c),
                                         ptr(lv_c), ptr(a_hat), ptr(g_loss), ptr(g_action), ptr(g_kl), ptr(g_a_hat), ptr(g_pad), ptr(d_hs),
                                         E, B * E, ptr(dWa), ptr(dba), ptr(dWp), ptr(dbp), ptr(dmu), ptr(dlv),
                                         current_stream()), "pcm_act_heads_loss_bwd")
        return (d_hs.transpose(0, 1), None if slots[0] is not None else dWa, None if slots[1] is not None else dba,
                None if (g_pad is None or pslots[0] is not None) else dWp, None if (g_pad is None or pslots[1] is not None) else dbp,
                None, None, dmu, dlv, None)


def act_heads_loss(hs, action_head, is_pad_head, actions, is_pad, mu, logvar, kl_weight, sig_start=None, n_pos=0, w_pos=1.0):
    """(a_hat, is_pad_hat, loss, action_loss, kl_loss) of the ACT heads on decoder features hs (B, Q, E) (any (b, q)
    strides, contiguous channels); `actions is None` -> (a_hat, is_pad_hat, None, None, None).  Outputs
    d >= sig_start pass through a sigmoid, the first n_pos dims of the squared error are weighted by w_pos (RLBench)."""
    _need_cuda(hs)
    A, E = action_head.weight.shape
    if E % 128 or E > 1024 or A + 1 > 16 or hs.dtype != torch.float32 or hs.stride(2) != 1 or hs.stride(0) % 4 or hs.stride(1) % 4:
        return None  # caller composes the heads from linear() (test fixtures with odd widths only)
    cfg = (A if sig_start is None else sig_start, n_pos, w_pos, kl_weight)
    return _ActHeadsLoss.apply(hs, action_head.weight, action_head.bias, is_pad_head.weight, is_pad_head.bias, actions, is_pad,
                               mu, logvar, cfg)


def clip_adamw_step(param, grad, exp_avg, exp_avg_sq, hyper, sumsq, norm_out, param_bf16=None, zero_grad=False):
    """Fused clip-by-global-norm + AdamW over flat fp32 buffers, in place (csrc/optimizer.cu).
    `hyper` = device tensor [lr, beta1, beta2, eps, wd, bias_corr1, bias_corr2, clip_norm, grad_scale]."""
    _need_cuda(param)
    K.clip_adamw_step(param, grad, exp_avg, exp_avg_sq, hyper, sumsq, norm_out, param_bf16, zero_grad)
    return norm_out


# ------------------------------------------------------------------------------------------------
# live kernel timing for bench.py's roofline (CUDA events on the launching stream, kernels.TIMER)
# ------------------------------------------------------------------------------------------------
KERNEL_TIMER = K.TIMER


def retime_gemm_shapes(kstats, iters=20):
    """Second, tighter live measurement for the roofline: every GEMM configuration the step launched (shape,
    operand majors, output type, split mode -- the tags the kernel timer recorded) is re-launched `iters` times
    BACK TO BACK (replayed from a CUDA graph, like the step itself) between one pair of CUDA events on the
    launching stream, on operands of the same shapes (two rotating sets).  Bracketing a single ~15 us launch with its own event pair adds several microseconds of launch
    gap per launch (the raw figure is kept as `total_ms_raw_events`); inside the CUDA-graph replay launches are
    back to back as they are here.  Returns total ms for one pass over all recorded launches."""
    st = kstats.get("gemm_tcgen05")
    if not st or "by_shape" not in st:
        return None
    total_ms, dev = 0.0, torch.device("cuda", torch.cuda.current_device())
    for tag, rec in st["by_shape"].items():
        t = eval(tag)  # tags are tuples written by kernels._Timer
        if len(t) != 8:  # grouped weight-gradient launches (dozens of problems each): long enough for their own event pair
            rec["isolated_us"] = max(rec["total_ms"] / rec["launches"] - st.get("event_pair_overhead_us", 0.0) * 1e-3, 0.0) * 1e3
            total_ms += rec["isolated_us"] * 1e-3 * rec["launches"]
            continue
        M, N, Kd, batch, a_mn, b_mn, dt, split_k = t
        sets = []
        for _ in range(2):
            a = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
            b = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
            out = torch.zeros((M, N), dtype=torch.bfloat16 if dt == "bfloat16" else torch.float32, device=dev)
            sets.append((a, b, out))
        acc = split_k == 0

        def launch(i):
            a, b, out = sets[i & 1]
            K.gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, accumulate=acc, split_k=split_k)

        was = K.TIMER.enabled
        K.TIMER.enabled = False
        for i in range(3):
            launch(i)
        torch.cuda.synchronize()
        # the Python / ctypes launch path costs more host time than the short GEMMs run: replay the `iters`
        # launches from a CUDA graph (as the step itself does) so the event pair brackets device time only
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(iters):
                launch(i)
        g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        K.TIMER.enabled = was
        per = e0.elapsed_time(e1) / iters
        rec["isolated_us"] = per * 1e3
        total_ms += per * rec["launches"]
    st["total_ms_isolated"] = total_ms
    return total_ms


def roofline_for(kstats, peaks, steps):
    """Roofline object for the dominant timed kernel family of the step (bench.py): the tcgen05 GEMM.
    achieved = algorithmic FLOPs (2*M*N*K per launch, summed) / CUDA-event time of those launches."""
    if not kstats or "gemm_tcgen05" not in kstats:
        return None
    st = kstats["gemm_tcgen05"]
    have = "bf16_tflops_sustained" in peaks
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    t_ms = st.get("total_ms_isolated") or st["total_ms"]
    achieved = st["flops"] / (t_ms * 1e-3) / 1e12
    traffic, traffic_note = None, None
    try:  # DRAM bytes per launch of the dominant shape from the committed `ncu --set full` capture
        import json
        from pathlib import Path

        t = json.loads((Path(__file__).resolve().parent.parent / "profiles" / "r2_gemm_traffic.json").read_text())
        traffic = t["dram_bytes_per_launch"]
        traffic_note = (f"dram__bytes_read+write per launch of the dominant shape M={t['shape']['M']} N={t['shape']['N']} "
                        f"K={t['shape']['K']} (algorithmic {t['algorithmic_bytes_per_launch']} B; output partly L2-resident), "
                        "profiles/r2_gemm_traffic.json")
    except Exception:
        pass
    return {"kernel": "gemm_tcgen05_kernel (all projection / attention / weight-gradient GEMMs of the step)",
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured"
                            if have else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md), of fallback"),
            "launches_per_step": st["launches"] / steps, "ms_per_step": t_ms / steps,
            "ms_per_step_event_pair_per_launch": st["total_ms"] / steps,
            "ms_per_step_raw_events": st.get("total_ms_raw_events", st["total_ms"]) / steps,
            "event_pair_overhead_us": st.get("event_pair_overhead_us", 0.0),
            "algorithmic_gflop_per_step": st["flops"] / steps / 1e9,
            "algorithmic_gbytes_per_step": st["bytes"] / steps / 1e9,
            "note": "launch list and algorithmic flops recorded on an eager (non-graph) replica of the step; duration of "
                    "each recorded GEMM configuration = 20 back-to-back launches (graph replay) between one CUDA-event pair on the launch "
                    "stream (ms_per_step); also given: one event pair per launch minus the calibrated empty-pair cost "
                    "(ms_per_step_event_pair_per_launch, inflated by the per-launch gap) and its raw sum; "
                    "traffic (dram bytes per launch from ncu --set full) is in profiles/ for the dominant shape"}
