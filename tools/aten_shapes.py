import sys
from collections import defaultdict
from pathlib import Path
import torch
from torch.profiler import ProfilerActivity, profile
ROOT = Path("/root/repo")
sys.path.insert(0, str(ROOT))
import bench
from pointcloudmatters_b200.act import build_policy
from pointcloudmatters_b200.bc_module import ACTBCModule
from pointcloudmatters_b200.data import synthetic_act_batch, to_device
dev = torch.device("cuda:0")
torch.manual_seed(0)
module = ACTBCModule(build_policy(bench.CFG2).to(dev).train(), total_steps=1000)
hb = synthetic_act_batch(64, 1024, seed=1)
b = to_device(hb, dev)
b["pcds"]["n_max"] = hb["pcds"]["n_max"]
for i in range(3):
    module.training_step(b, i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    module.training_step(b, 3)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    dt = getattr(e, "self_device_time_total", 0)
    if e.key.startswith("aten::") and dt > 0:
        rows.append((dt, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
for dt, c, k, s in rows[:60]:
    print(f"{dt:9.1f} us {c:4d}x {k:24s} {s}")
