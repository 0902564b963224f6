"""Debug: determinism of the DP training step, eager vs eager vs graph (losses over 6 steps)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pointcloudmatters_b200.bc_module import DiffusionPolicyBCModule  # noqa: E402
from pointcloudmatters_b200.data import synthetic_dp_batch, to_device  # noqa: E402
from pointcloudmatters_b200.diffusion import build_dp_policy  # noqa: E402

cfg = dict(qpos_dim=9, action_dim=7, backbone_classes=32, n_obs_steps=2, pcd_nsample=16, pcd_npoints=64,
           pcd_hidden_dim=32, projector_layers=1, projector_channels=[32, 64, 64], horizon=16,
           diffusion_step_embed_dim=64, down_dims=[64, 128], kernel_size=5, n_groups=8, goal_dim=0)


def run(graph, nb=4):
    torch.manual_seed(1)
    policy = build_dp_policy(cfg).cuda().train()
    policy.normalizer.set_identity({"qpos": 9, "action": 7}).cuda()
    module = DiffusionPolicyBCModule(policy, total_steps=50, use_cuda_graph=graph)
    out, norms = [], []
    for step in range(6):
        batch = synthetic_dp_batch(nb, 128, seed=500 + step)
        gen = torch.Generator().manual_seed(step)
        gb = to_device(batch, "cuda")
        gb["obs"]["pcds"]["n_max"] = batch["obs"]["pcds"]["n_max"]
        gb["_noise"] = torch.randn(nb, 16, 7, generator=gen).cuda()
        gb["_timesteps"] = torch.randint(0, 100, (nb,), generator=gen).cuda()
        out.append(float(module.training_step(gb, step)))
        norms.append(float(module._trainer.last_grad_norm))
    return out, norms


for nb in (4, 16):
    print("batch", nb)
    for name, g in (("eager A", False), ("eager B", False), ("graph  ", True)):
        o, n = run(g, nb)
        print(name, ["%.7f" % v for v in o], ["%.5f" % v for v in n])
