"""Host-side cost of one cfg-2 training_step call with a pinned host batch (CUDA-graph path): wall time from call to
return with the device idle at call time = what the e2e figure pays on top of the device time of the step, because the
per-step loss read-back drains the stream.  usage: python tools/host_overhead.py [steps]   (prints a cProfile top list)"""
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pointcloudmatters_b200.act import build_policy  # noqa: E402
from pointcloudmatters_b200.bc_module import ACTBCModule  # noqa: E402
from pointcloudmatters_b200.data import synthetic_act_batch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
torch.manual_seed(0)
module = ACTBCModule(build_policy(bench.CFG2).cuda().train(), total_steps=1000, use_cuda_graph=True)
hosts = [synthetic_act_batch(64, 1024, seed=s, pin=True) for s in (1, 2)]
for i in range(8):
    float(module.training_step(hosts[i % 2], i))
torch.cuda.synchronize()
call, total = [], []
for i in range(steps):
    t0 = time.perf_counter()
    loss = module.training_step(hosts[i % 2], i)
    t1 = time.perf_counter()
    float(loss)
    t2 = time.perf_counter()
    call.append((t1 - t0) * 1e6)
    total.append((t2 - t0) * 1e6)
call.sort(); total.sort()
print(f"training_step call (enqueue) median {call[len(call) // 2]:.0f} us; call + loss read-back median {total[len(total) // 2]:.0f} us")
pr = cProfile.Profile()
pr.enable()
for i in range(steps):
    float(module.training_step(hosts[i % 2], i))
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
