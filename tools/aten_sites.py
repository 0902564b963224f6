#!/usr/bin/env python
"""Where do the remaining ATen kernels of the cfg-2 step come from?  One eager step under torch.profiler with Python
stacks; prints every aten op that launched device work, with its device time and the innermost frames inside this
repository.   python tools/aten_sites.py > gpurun_out/aten_sites.txt"""
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pointcloudmatters_b200.act import build_policy  # noqa: E402
from pointcloudmatters_b200.bc_module import ACTBCModule  # noqa: E402
from pointcloudmatters_b200.data import synthetic_act_batch, to_device  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
module = ACTBCModule(build_policy(bench.CFG2).to(dev).train(), total_steps=1000)
hb = synthetic_act_batch(64, 1024, seed=1)
b = to_device(hb, dev)
b["pcds"]["n_max"] = hb["pcds"]["n_max"]
for i in range(3):
    module.training_step(b, i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    module.training_step(b, 3)
    torch.cuda.synchronize()
agg = defaultdict(lambda: [0, 0.0])
for e in prof.key_averages(group_by_stack_n=12):
    dt = getattr(e, "self_device_time_total", 0)
    if not e.key.startswith("aten::") or dt <= 0:
        continue
    frames = [f for f in e.stack if "pointcloudmatters_b200" in f or "/repo/" in f]
    site = " <- ".join(f.split("/")[-1] for f in frames[:3]) or "(autograd engine / no repo frame)"
    k = (e.key, site)
    agg[k][0] += e.count
    agg[k][1] += dt
tot = sum(v[1] for v in agg.values())
print(f"# aten ops with device time, one eager cfg-2 step: {tot:.0f} us in {sum(v[0] for v in agg.values())} calls")
for (op, site), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t:8.1f} us {n:4d}x  {op:28s} {site}")

# second view: by (op, input shapes, enclosing non-aten profiler range = the autograd node / python function that issued it)
agg2 = defaultdict(lambda: [0, 0.0])
for e in prof.events():
    dt = getattr(e, "self_device_time_total", 0)
    if not e.name.startswith("aten::") or dt <= 0:
        continue
    par, owner = e.cpu_parent, "(top level)"
    while par is not None:
        if not par.name.startswith("aten::"):
            owner = par.name
            break
        par = par.cpu_parent
    shapes = str([s for s in (e.input_shapes or []) if s])[:70]
    agg2[(e.name, shapes, owner[:70])][0] += 1
    agg2[(e.name, shapes, owner[:70])][1] += dt
print("\n# by op / input shapes / issuing range")
for (op, shp, owner), (n, t) in sorted(agg2.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{t:8.1f} us {n:4d}x  {op:18s} {shp:72s} {owner}")
