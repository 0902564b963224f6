#!/usr/bin/env python
"""Per-stage device times of the cfg-2 training step (SURVEY.md 8d "per-stage ms"): CUDA events at the stage markers
(functional.stage -- the same places carry NVTX ranges under PCM_NVTX=1) on an EAGER step.  Stages launched on side streams
overlap the main stream, so the rows do not add up to the step; nested rows (encoder / decoder inside "transformer + heads
+ loss", everything inside "forward") are indented.  The graph-replayed step of bench.py is ~5 % shorter than the eager
sum (no host launch gaps).

    python tools/stage_times.py [steps] > profiles/rN_stage_times.txt
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pointcloudmatters_b200 import functional as PF  # noqa: E402
from pointcloudmatters_b200.act import build_policy  # noqa: E402
from pointcloudmatters_b200.bc_module import ACTBCModule  # noqa: E402
from pointcloudmatters_b200.data import synthetic_act_batch, to_device  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
torch.manual_seed(0)
module = ACTBCModule(build_policy(bench.CFG2).to(dev).train(), total_steps=1000)
hb = synthetic_act_batch(64, 1024, seed=1)
b = to_device(hb, dev)
b["pcds"]["n_max"] = hb["pcds"]["n_max"]
for i in range(4):
    module.training_step(b, i)
torch.cuda.synchronize()
PF.STAGE_EVENTS = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    torch.cuda._sleep(int(0.02 * 1.9e9))  # let the host run ahead: events then bracket device time, not launch latency
    module.training_step(b, i)
e1.record()
torch.cuda.synchronize()
agg = {}
for name, a, c in PF.STAGE_EVENTS:
    agg.setdefault(name, []).append(a.elapsed_time(c))
PF.STAGE_EVENTS = None
nest = {"fps+knn+sine (side stream)": 1, "cvae encoder (side stream)": 1, "cvae encoder": 1, "pointnet + set abstraction": 1,
        "transformer + heads + loss": 1}
print(f"# cfg-2 eager step, {steps} steps, one B200; median ms per stage (CUDA events on the stage's stream)")
for name, v in agg.items():
    v = sorted(v)
    depth = 2 if name.startswith(("encoder x", "decoder x")) else nest.get(name, 0)
    print(f"{'  ' * depth}{name:45s} {v[len(v) // 2]:8.3f} ms")
print(f"total (incl. the 20 ms spin per step) {e0.elapsed_time(e1) / steps:8.3f} ms/step")
