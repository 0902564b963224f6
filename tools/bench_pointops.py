"""Micro-benchmark (CUDA events) of the sm_100a FPS / kNN kernels vs the unmodified reference
kernels (oracle/_ref) at BASELINE cfg-2 / cfg-4 shapes.  Usage: python tools/bench_pointops.py"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pointcloudmatters_b200 import pointops as P  # noqa: E402
from pointcloudmatters_b200._lib import lib  # noqa: E402
from tests import _ref  # noqa: E402
from tests._data import clouds  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


out = {}
for b, n, m in [(64, 1024, 512), (32, 4096, 2048), (8, 4096, 2048)]:
    xyz, off, noff = clouds(b, n, m, seed=1)
    t_xyz, t_off, t_noff = [torch.from_numpy(a).cuda() for a in (xyz, off, noff)]
    key = f"b{b}_n{n}_m{m}"
    res = {}
    for T in (0, 128, 256, 512, 1024):
        if T and (n + T - 1) // T > 8:
            continue
        lib.pcm_tune_fps_threads(T)
        res[f"fps_T{T}_ms"] = timeit(lambda: P.farthest_point_sampling(t_xyz, t_off, t_noff, n_max=n, m_total=b * m))
    lib.pcm_tune_fps_threads(0)
    fps = P.farthest_point_sampling(t_xyz, t_off, t_noff, n_max=n, m_total=b * m)
    q = t_xyz[fps.long()].contiguous()
    res["knn16_ms"] = timeit(lambda: P.knn_query(16, t_xyz, t_off, q, t_noff))
    if _ref.available():
        res["ref_fps_ms"] = timeit(lambda: _ref.farthest_point_sampling(t_xyz, t_off, t_noff), iters=5, warm=1)
        res["ref_knn16_ms"] = timeit(lambda: _ref.knn_query(16, t_xyz, t_off, q, t_noff), iters=5, warm=1)
    out[key] = res
    print(key, json.dumps(res))
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bench_pointops.json").write_text(json.dumps(out, indent=1))
