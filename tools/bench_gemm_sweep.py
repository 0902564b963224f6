"""Tile-shape / split-K sweep of pcm_gemm_bf16 at the step's latency-bound shapes (decoder, FFN-32)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200._lib import lib  # noqa: E402
from pointcloudmatters_b200.kernels import gemm_bf16  # noqa: E402


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


shapes = [
    # M, N, K, a_mn, b_mn, out dtype, split list, name
    (6400, 512, 512, 0, 0, torch.bfloat16, [1], "dec proj bf16"),
    (6400, 512, 512, 0, 0, torch.float32, [1], "dec proj f32"),
    (6400, 512, 512, 0, 1, torch.float32, [1], "dec dX"),
    (512, 512, 6400, 1, 1, torch.float32, [1, 2, 4, 6, 12], "dec dW"),
    (512, 512, 32960, 1, 1, torch.float32, [4, 9, 18], "enc dW"),
    (32960, 512, 512, 0, 0, torch.bfloat16, [1], "enc proj bf16"),
    (6400, 32, 512, 0, 0, torch.float32, [1], "ffn1 dec"),
    (6400, 512, 32, 0, 0, torch.float32, [1], "ffn2 dec"),
    (32, 512, 6400, 1, 1, torch.float32, [1, 4, 12], "ffn dW1 dec"),
    (512, 32, 6400, 1, 1, torch.float32, [1, 4, 12], "ffn dW2 dec"),
    (32960, 32, 512, 0, 0, torch.float32, [1], "ffn1 enc"),
    (32960, 512, 32, 0, 0, torch.float32, [1], "ffn2 enc"),
    (32, 512, 32960, 1, 1, torch.float32, [4, 12, 37], "ffn dW1 enc"),
]
for (M, N, K, a_mn, b_mn, odt, splits, name) in shapes:
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda", dtype=odt)
    res = []
    for bn in (64, 128, 256):
        if bn > 64 and N <= 64:
            continue
        lib.pcm_gemm_debug_force_bn(bn)
        for sk in splits:
            if sk > 1 or (a_mn and b_mn):
                t = timeit(lambda: gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, accumulate=True, split_k=sk))
            else:
                t = timeit(lambda: gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out))
            res.append(f"bn{bn}/sk{sk}: {t:6.1f}us")
    lib.pcm_gemm_debug_force_bn(0)
    print(f"{name:16s} M{M} N{N} K{K}: " + "  ".join(res), flush=True)
