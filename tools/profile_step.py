"""Per-kernel breakdown of the cfg-2 (ACT) or cfg-3-shape (Diffusion Policy) training step
(torch.profiler, CUDA activities).
Usage: python tools/profile_step.py [steps] [act|dp] [batch] -> gpurun_out/profile_step[_dp].txt, kernels_per_step[_dp].txt"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pointcloudmatters_b200.act import build_policy  # noqa: E402
from pointcloudmatters_b200.bc_module import ACTBCModule  # noqa: E402
from pointcloudmatters_b200.data import synthetic_act_batch, to_device  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
which = sys.argv[2] if len(sys.argv) > 2 else "act"
SUFFIX = "" if which == "act" else "_" + which
dev = torch.device("cuda:0")
torch.manual_seed(0)
if which == "act":
    policy = build_policy(bench.CFG2).to(dev).train()
    module = ACTBCModule(policy, total_steps=1000)
    hb = synthetic_act_batch(int(sys.argv[3]) if len(sys.argv) > 3 else 64, 1024, seed=1)
    b = to_device(hb, dev)
    b["pcds"]["n_max"] = hb["pcds"]["n_max"]
else:
    from pointcloudmatters_b200.bc_module import DiffusionPolicyBCModule
    from pointcloudmatters_b200.data import synthetic_dp_batch
    from pointcloudmatters_b200.diffusion import DP_MODEL_CFG, build_dp_policy

    cfg = dict(DP_MODEL_CFG, pcd_npoints=512)
    policy = build_dp_policy(cfg).to(dev).train()
    policy.normalizer.set_identity({"qpos": 9, "action": 7}).to(dev)
    module = DiffusionPolicyBCModule(policy, total_steps=1000)
    hb = synthetic_dp_batch(int(sys.argv[3]) if len(sys.argv) > 3 else 16, 1024, seed=1)
    b = to_device(hb, dev)
    b["obs"]["pcds"]["n_max"] = hb["obs"]["pcds"]["n_max"]
    SUFFIX += "_b" + str(hb["action"].shape[0])
for i in range(3):
    module.training_step(b, i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(steps):
        module.training_step(b, i)
    torch.cuda.synchronize()
tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
# every device kernel (name, launches / step, us / step), sorted by time
kern = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        d = kern.setdefault(e.name, [0, 0.0])
        d[0] += 1
        d[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
lines = [f"{v[1] / steps:10.1f} us/step {v[0] / steps:7.1f} launches/step  {k[:150]}" for k, v in sorted(kern.items(), key=lambda kv: -kv[1][1])]
total = sum(v[1] for v in kern.values()) / steps
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"kernels_per_step{SUFFIX}.txt").write_text(f"# total {total:.1f} us/step over {sum(v[0] for v in kern.values()) / steps:.0f} launches/step (torch.profiler, eager step)\n" + "\n".join(lines) + "\n")
out = ROOT / "gpurun_out" / f"profile_step{SUFFIX}.txt"
out.parent.mkdir(exist_ok=True)
out.write_text(tab)
print(tab[-9000:])

# wall vs GPU-busy
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(steps):
    module.training_step(b, i)
t_cpu_issue = time.perf_counter() - t0
torch.cuda.synchronize(); t1 = time.perf_counter() - t0
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
gpu_ms = sum(e.cuda_time for e in evs) / 1e3 if evs else float("nan")
ka = prof.key_averages()
self_cuda = sum(getattr(k, "self_device_time_total", getattr(k, "self_cuda_time_total", 0)) for k in ka) / 1e3
summary = f"steps={steps} wall_ms_per_step={t1/steps*1e3:.2f} cpu_issue_ms_per_step={t_cpu_issue/steps*1e3:.2f} gpu_kernel_ms_per_step={self_cuda/steps:.2f} kernels_per_step={sum(k.count for k in ka if k.device_type == torch.autograd.DeviceType.CUDA)/steps:.0f}"
print(summary)
(out.parent / f"profile_summary{SUFFIX}.txt").write_text(summary + "\n")
