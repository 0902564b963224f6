"""Micro-benchmark of the fused attention kernels at the cfg-2 shapes (B = 64, 8 heads, head dim 64, dropout 0.1):
main-encoder self-attention (L = S = 515), decoder cross-attention (L = 100, S = 515), decoder self-attention (L = S = 100).
usage: python tools/flash_micro.py [iters] [fwd|bwd|both]     CUDA events around one launch, L2 flushed before each,
median; the backward time covers its three launches (delta, main kernel, dQ store)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200 import kernels as K  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2] if len(sys.argv) > 2 else "both"
B, nh = 64, 8
Z, E = B * nh, nh * 64
sb = torch.tensor([1234567], dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def med(fn):
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for L, S in ((515, 515), (100, 515), (100, 100)):
    q = torch.randn(Z * L, 64, device="cuda").bfloat16()
    k = torch.randn(Z * S, 64, device="cuda").bfloat16()
    v = torch.randn(Z * S, 64, device="cuda").bfloat16()
    do = torch.randn(Z * L, 64, device="cuda").bfloat16()
    buf = torch.empty(L * B, E, dtype=torch.bfloat16, device="cuda")
    kvb = torch.empty(S * B, 2 * E, dtype=torch.bfloat16, device="cuda")
    O, lse = K.flash_attn_fwd(q, k, v, B, nh, L, S, None, 0.125, 0.1, sb, 77)
    out = [f"L={L} S={S}:"]
    if which in ("fwd", "both"):
        tf = med(lambda: K.flash_attn_fwd(q, k, v, B, nh, L, S, None, 0.125, 0.1, sb, 77))
        out.append(f"fwd {tf:.1f} us ({4.0 * Z * L * S * 64 / tf * 1e-6:.0f} TFLOP/s)")
    if which in ("bwd", "both"):
        tb = med(lambda: K.flash_attn_bwd(q, k, v, O, do, lse, B, nh, L, S, None, 0.125, 0.1, sb, 77, buf, kvb[:, :E], kvb[:, E:]))
        out.append(f"bwd {tb:.1f} us ({10.0 * Z * L * S * 64 / tb * 1e-6:.0f} TFLOP/s)")
    print(" ".join(out), flush=True)
