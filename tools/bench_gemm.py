"""GEMM micro-benchmark: pcm_gemm_bf16 (tcgen05) vs torch/cuBLAS bf16 at the policy's shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200.kernels import gemm_bf16  # noqa: E402


def timeit(fn, iters=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for (M, N, K, a_mn, b_mn, sk, name) in [
    (32960, 1024, 512, 0, 0, 1, "enc QK proj fwd"), (32960, 512, 512, 0, 0, 1, "enc V/out proj fwd"),
    (65536, 512, 512, 0, 0, 1, "SA Pf = feat Wf^T"), (65536, 512, 128, 0, 0, 1, "pointnet conv5"),
    (32960, 512, 1024, 0, 1, 1, "dX = dY W (QK)"), (1024, 512, 32960, 1, 1, 32, "dW = dY^T X (QK)"),
    (512, 512, 65536, 1, 1, 64, "dW SA"), (6400, 512, 512, 0, 0, 1, "dec proj"),
]:
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    out = torch.zeros(M, N, device="cuda")
    if sk > 1:
        t = timeit(lambda: gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out, accumulate=True, split_k=sk))
    else:
        t = timeit(lambda: gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), out=out))
    A = a.t() if a_mn else a
    B = b if b_mn else b.t()
    t_ref = timeit(lambda: torch.matmul(A, B))
    fl = 2.0 * M * N * K
    print(f"{name:22s} M{M} N{N} K{K} mn({a_mn},{b_mn}) sk{sk}: pcm {t*1e3:8.1f} us {fl/t/1e9:7.1f} TF/s | cublas(bf16 out) {t_ref*1e3:8.1f} us {fl/t_ref/1e9:7.1f} TF/s")
