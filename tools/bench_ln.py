"""Launch-shape sweep of the LayerNorm backward and colsum kernels (both end in per-column global atomics).
usage: python tools/bench_ln.py  -> JSON lines: avg us per launch for each (rows, knob) pair, CUDA events, 40 launches
over 4 rotating buffer sets (inputs of the 32 960-row case exceed L2 together; the 6 400-row ones are L2-resident in the
step as well: they are produced by the preceding kernel)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200 import kernels as K  # noqa: E402
from pointcloudmatters_b200._lib import lib  # noqa: E402


def timeit(fn, n=40):
    """avg us per launch of `fn(i)`, i = 0..n-1, replayed from a CUDA graph (pure device time: the Python wrappers
    cost more host time than these kernels run)."""
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3


def main():
    C = 512
    sb = torch.tensor([1234567], dtype=torch.int64, device="cuda")
    for rows in (6400, 6528, 32960):
        sets = []
        for _ in range(4):
            dy, h = torch.randn(rows, C, device="cuda"), torch.randn(rows, C, device="cuda")
            mean, rstd = torch.randn(rows, device="cuda"), torch.rand(rows, device="cuda") + 0.5
            sets.append((dy, h, mean, rstd, dy.bfloat16()))
        gamma = torch.randn(C, device="cuda")
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        out = torch.zeros(C, device="cuda")
        xs = [torch.randn(rows, C, device="cuda") for _ in range(4)]
        beta = torch.randn(C, device="cuda")
        pos = torch.randn(rows // 64, C, device="cuda") if rows % 64 == 0 else None
        for cap in (148, 296, 592, 1184):
            lib.pcm_ln_debug_tune(cap, 0, 0, 0)
            us = timeit(lambda i: K.add_dropout_ln_fwd(xs[i % 4], sets[i % 4][1], gamma, beta, 1e-5, 0.1, sb, 7, want_bf16=True,
                                                       pos=pos, pos_row_div=64 if pos is not None else 1))
            print(json.dumps({"kernel": "ln_fwd", "rows": rows, "max_ctas": cap, "us": round(us, 2)}))
        lib.pcm_ln_debug_tune(1184, 0, 0, 0)
        for cap in (222, 296):
            lib.pcm_ln_debug_tune(0, cap, 0, 0)
            us = timeit(lambda i: K.add_dropout_ln_bwd(sets[i % 4][0], sets[i % 4][1], sets[i % 4][2], sets[i % 4][3], gamma,
                                                       0.1, sb, 7, True, dg, db, True))
            print(json.dumps({"kernel": "ln_bwd", "rows": rows, "max_ctas": cap, "us": round(us, 2)}))
        for ctas in (592,):
            for min_rows in (32,):
                lib.pcm_ln_debug_tune(0, 0, ctas, min_rows)
                us = timeit(lambda i: K.colsum(sets[i % 4][4], out))
                print(json.dumps({"kernel": "colsum_bf16", "rows": rows, "ctas": ctas, "min_rows": min_rows, "us": round(us, 2)}))


if __name__ == "__main__":
    main()
