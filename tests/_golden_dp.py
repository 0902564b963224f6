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


GOLDEN_DPENC = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dpenc_*.npz")))


def load_encoder(path):
    """Encoder-only fixture (reference PCDObsEncoder variants, oracle/gen_golden_dp.py `encoder`)."""
    g = np.load(path)
    cfg = {k: ast.literal_eval(v) for k, v in zip(g["meta/cfg_keys"].tolist(), g["meta/cfg_vals"].tolist())}
    state = {k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    obs = {"pcds": {k[len("in/pcds/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in/pcds/")},
           "qpos": torch.from_numpy(g["in/qpos"])}
    grads = {k[len("grad/"):]: g[k] for k in g.files if k.startswith("grad/")}
    post = {k[len("post/"):]: g[k] for k in g.files if k.startswith("post/")}
    return cfg, state, obs, torch.from_numpy(g["out/features"]), torch.from_numpy(g["in/probe"]), grads, post


def encoder_kwargs(cfg):
    sm = {"obs": {"pcds": {"shape": [6], "type": "pcd"}, "qpos": {"shape": [cfg["qpos_dim"]], "type": "low_dim"}}}
    return sm, dict(share_pcd_model=True, n_obs_step=2, pcd_nsample=cfg["pcd_nsample"], pcd_npoints=cfg["pcd_npoints"],
                    use_mask=cfg["use_mask"], bg_ratio=cfg["bg_ratio"], pcd_hidden_dim=cfg["pcd_hidden_dim"],
                    projector_layers=cfg["projector_layers"], projector_channels=cfg["projector_channels"],
                    pre_sample=cfg["pre_sample"], in_channel=6)
