"""ncu target: a few EAGER cfg-2 training steps (no CUDA graph), for a launch list of the step:
    ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1150 --launch-count 1300 --csv \
        --log-file gpurun_out/launches.csv python tools/ncu_step_target.py
(the window must contain two fused-AdamW launches; tools/summarize_launches.py counts what lies between them)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from pointcloudmatters_b200.act import build_policy  # noqa: E402
from pointcloudmatters_b200.bc_module import ACTBCModule  # noqa: E402
from pointcloudmatters_b200.data import synthetic_act_batch, to_device  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)
module = ACTBCModule(build_policy(bench.CFG2).cuda().train(), total_steps=1000, use_cuda_graph=False)
hb = synthetic_act_batch(64, 1024, seed=1)
b = to_device(hb, "cuda")
b["pcds"]["n_max"] = hb["pcds"]["n_max"]
for i in range(steps):
    module.training_step(b, i)
torch.cuda.synchronize()
print("done", steps)
