"""Loader for the Diffusion-Policy fixtures written by oracle/gen_golden_dp.py (reference modules, CPU fp32)."""
import ast
import glob
import os

import numpy as np
import torch

GOLDEN_DP = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dp_*.npz")))


def load(path):
    g = np.load(path)
    cfg = {k: ast.literal_eval(v) for k, v in zip(g["meta/cfg_keys"].tolist(), g["meta/cfg_vals"].tolist())}
    state = {k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    batch = {"obs": {"qpos": torch.from_numpy(g["in/obs/qpos"]),
                     "pcds": {k: torch.from_numpy(g["in/obs/pcds/" + k]) for k in ("coord", "grid_coord", "feat", "offset")}},
             "action": torch.from_numpy(g["in/action"]),
             "_noise": torch.from_numpy(g["in/noise"]), "_timesteps": torch.from_numpy(g["in/timesteps"])}
    if "in/goal/task_emb" in g.files:
        batch["goal"] = {"task_emb": torch.from_numpy(g["in/goal/task_emb"])}
    grads = {k[len("grad/"):]: g[k] for k in g.files if k.startswith("grad/")}
    post = {k[len("post/"):]: g[k] for k in g.files if k.startswith("post/")}
    return cfg, state, batch, float(g["out/loss"]), grads, post, g["meta/nograd"].tolist()


def load_prediction(path):
    """Inference fixture: the reference `predict_action` (eval mode, 10 DDPM steps) on the same observations /
    initial state, with every noise draw of the sampling loop recorded in order."""
    g = np.load(path)
    return ([torch.from_numpy(x) for x in g["pred/noises"]], torch.from_numpy(g["pred/action"]),
            torch.from_numpy(g["pred/action_pred"]))
