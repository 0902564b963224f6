"""Micro-benchmark of the dim_feedforward = 32 FFN: the fused kernels (csrc/ffn_fused.cu) against the three-launch path
(GEMM -> dropout -> GEMM) at the two row counts of BASELINE cfg-2 (main encoder 64 x 515 rows, decoder 64 x 100 rows).
usage: python tools/ffn_micro.py [iters]          (CUDA events on the launching stream, L2 flushed between launches)
       python tools/ffn_micro.py ncu              (two launches of each fused kernel per size, for an ncu capture)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200 import kernels as K  # noqa: E402
from pointcloudmatters_b200._lib import check, current_stream, lib, ptr  # noqa: E402

E, Hd, P = 512, 32, 0.1
mode = sys.argv[1] if len(sys.argv) > 1 else "20"
dev = "cuda"
w1 = (torch.randn(Hd, E, device=dev) * 0.05).bfloat16()
w2 = (torch.randn(E, Hd, device=dev) * 0.05).bfloat16()
b1, b2 = torch.randn(Hd, device=dev) * 0.1, torch.randn(E, device=dev) * 0.1
sb = torch.tensor([1234567], dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def fused_fwd(x, hd, y):
    check(lib.pcm_ffn32_fwd(x.shape[0], E, Hd, ptr(x), x.stride(0), ptr(w1), ptr(b1), ptr(w2), ptr(b2), P, ptr(sb), 7, ptr(hd), ptr(y),
                            current_stream()), "fwd")


def fused_bwd(dy, hd, dh, dx):
    check(lib.pcm_ffn32_bwd(dy.shape[0], E, Hd, ptr(dy), dy.stride(0), ptr(hd), ptr(w1), ptr(w2), P, ptr(dh), ptr(dx), current_stream()),
          "bwd")


def three_fwd(x, hd, y):
    h = K.gemm_bf16(x, w1, bias=b1, relu=True, out_dtype=torch.bfloat16)
    check(lib.pcm_ffn_dropout_fwd(h.shape[0], Hd, ptr(h), P, ptr(sb), 7, ptr(hd), current_stream()), "drop")
    K.gemm_bf16(hd, w2, bias=b2, out=y)


def three_bwd(dy, hd, dh, dx):
    dhd = K.gemm_bf16(dy, w2, b_mn=True)
    check(lib.pcm_ffn_relu_dropout_bwd_ex(dy.shape[0], Hd, ptr(dhd), ptr(hd), P, ptr(sb), 7, ptr(dh), None, current_stream()), "gate")
    K.gemm_bf16(dh, w1, b_mn=True, out=dx)


def timeit(fn, args, iters):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(*args)
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for rows in (32960, 6400):
    x = torch.randn(rows, E, device=dev).bfloat16()
    dy = torch.randn(rows, E, device=dev).bfloat16()
    hd = torch.empty(rows, Hd, device=dev, dtype=torch.bfloat16)
    dh = torch.empty_like(hd)
    y = torch.empty(rows, E, device=dev)
    dx = torch.empty(rows, E, device=dev)
    if mode == "ncu":
        for _ in range(2):
            fused_fwd(x, hd, y)
            fused_bwd(dy, hd, dh, dx)
        torch.cuda.synchronize()
        continue
    it = int(mode)
    mb = rows * E * (2 + 4) / 1e6
    f, b = timeit(fused_fwd, (x, hd, y), it), timeit(fused_bwd, (dy, hd, dh, dx), it)
    f3, b3 = timeit(three_fwd, (x, hd, y), it), timeit(three_bwd, (dy, hd, dh, dx), it)
    print(f"rows {rows}: fused fwd {f:.1f} us ({mb / f:.2f} TB/s)  bwd {b:.1f} us ({mb / b:.2f} TB/s) | "
          f"three-launch fwd {f3:.1f} us  bwd {b3:.1f} us", flush=True)
