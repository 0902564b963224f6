"""Helpers to load tests/golden/act_*.npz (made by oracle/gen_golden_act.py from the REFERENCE modules)."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

GOLDEN_ACT = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "act_*.npz")))


def load(path):
    g = np.load(path)
    cfg = {k: v for k, v in zip(g["meta/cfg_keys"].tolist(), g["meta/cfg_vals"].tolist())}
    for k in list(cfg):
        if k not in ("dropout", "kl_weight", "position_loss_weight", "bg_ratio"):
            cfg[k] = int(cfg[k])
    cfg["collision"] = bool(cfg.get("collision", 0))
    cfg["pre_sample"], cfg["use_mask"] = bool(cfg.get("pre_sample", 0)), bool(cfg.get("use_mask", 0))
    state = {k[len("state/"):]: torch.from_numpy(g[k].astype(np.float32) if g[k].dtype == np.float16 else g[k])
             for k in g.files if k.startswith("state/")}
    batch = {
        "pcds": {k: torch.from_numpy(g["in/pcds/" + k]) for k in ("coord", "grid_coord", "feat", "offset")},
        "qpos": torch.from_numpy(g["in/qpos"]), "actions": torch.from_numpy(g["in/actions"]),
        "is_pad": torch.from_numpy(g["in/is_pad"]), "goal_cond": torch.from_numpy(g["in/goal_cond"]),
        "_eps": torch.from_numpy(g["in/eps"]),
    }
    if "in/pcds/mask" in g.files:
        batch["pcds"]["mask"] = torch.from_numpy(g["in/pcds/mask"])
    out = {k[len("out/"):]: g[k] for k in g.files if k.startswith("out/")}
    grads = {k[len("grad/"):]: g[k] for k in g.files if k.startswith("grad/")}
    post = {k[len("post/"):]: g[k] for k in g.files if k.startswith("post/")}
    return cfg, state, batch, out, grads, post, g["meta/nograd"].tolist(), "rlbench" in os.path.basename(path)


def grad_summary(t):
    f = t.detach().double().flatten().cpu()
    step = max(1, f.numel() // 16)
    return np.concatenate([[f.norm().item(), f.sum().item()], f[::step][:16].numpy()])
