"""Synthetic batches in the reference's collate format (`pcd_collate_fn`,
src/utils/sparse_tensor_utils.py:65-82; dataset fields maniskill2_single_task_pcd_act.py:269-275):

    {pcds: {coord (sumN,3) f32, grid_coord (sumN,3) i64, feat (sumN,6) f32 = [rgb/127.5-1, xyz],
            offset (B) i64 cumulative}, qpos (B,Q) f32, actions (B,T,A) f32, is_pad (B,T) bool,
     goal_cond (B,G) f32}

Shapes / distributions follow SURVEY.md section 8d (coord ~ U([-0.5,0.5]^3), seed = 1000 + rank).
"""
from __future__ import annotations

import torch


def synthetic_act_batch(batch_size, n_points, *, num_queries=100, action_dim=7, qpos_dim=9, goal_cond_dim=3,
                        seed=1000, ragged=False, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    sizes = (torch.randint(int(0.75 * n_points), n_points + 1, (batch_size,), generator=g) if ragged
             else torch.full((batch_size,), n_points, dtype=torch.int64))
    total = int(sizes.sum())
    coord = torch.rand(total, 3, generator=g) - 0.5
    color = torch.randint(0, 256, (total, 3), generator=g).float() / 127.5 - 1.0
    grid = torch.floor(coord / 0.005).long()
    grid = grid - grid.min(0).values
    npad = torch.randint(0, num_queries // 2 + 1, (batch_size,), generator=g)
    is_pad = torch.arange(num_queries)[None, :] >= (num_queries - npad)[:, None]
    batch = {
        "pcds": {"coord": coord, "grid_coord": grid, "feat": torch.cat([color, coord], dim=1),
                 "offset": torch.cumsum(sizes, 0)},
        "qpos": torch.randn(batch_size, qpos_dim, generator=g),
        "actions": torch.randn(batch_size, num_queries, action_dim, generator=g),
        "is_pad": is_pad,
    }
    if goal_cond_dim > 0:
        batch["goal_cond"] = torch.randn(batch_size, goal_cond_dim, generator=g)
    if pin:
        batch = map_tensors(batch, lambda t: t.pin_memory())
    if device != "cpu":
        batch = to_device(batch, device)
    batch["pcds"]["n_max"] = int(sizes.max())  # host-known largest cloud: keeps the step sync-free
    return batch


def synthetic_dp_batch(batch_size, n_points, *, n_obs_steps=2, horizon=16, action_dim=7, qpos_dim=9, goal_dim=0,
                       seed=1000, ragged=False, device="cpu", pin=False):
    """Diffusion-Policy batch (maniskill2_single_task_pcd_dp.py:135-224): B * n_obs_steps clouds,
    `{obs: {qpos (B, horizon, Q), pcds: {...}}, action (B, horizon, A)[, goal: {task_emb (B, G)}]}`."""
    g = torch.Generator().manual_seed(seed)
    clouds = batch_size * n_obs_steps
    sizes = (torch.randint(int(0.75 * n_points), n_points + 1, (clouds,), generator=g) if ragged
             else torch.full((clouds,), n_points, dtype=torch.int64))
    total = int(sizes.sum())
    coord = torch.rand(total, 3, generator=g) - 0.5
    color = torch.randint(0, 256, (total, 3), generator=g).float() / 127.5 - 1.0
    grid = torch.floor(coord / 0.005).long()
    grid = grid - grid.min(0).values
    batch = {"obs": {"qpos": torch.randn(batch_size, horizon, qpos_dim, generator=g),
                     "pcds": {"coord": coord, "grid_coord": grid, "feat": torch.cat([color, coord], dim=1),
                              "offset": torch.cumsum(sizes, 0)}},
             "action": torch.randn(batch_size, horizon, action_dim, generator=g)}
    if goal_dim > 0:
        batch["goal"] = {"task_emb": torch.randn(batch_size, goal_dim, generator=g)}
    if pin:
        batch = map_tensors(batch, lambda t: t.pin_memory())
    if device != "cpu":
        batch = to_device(batch, device)
    batch["obs"]["pcds"]["n_max"] = int(sizes.max())
    return batch


def map_tensors(batch, fn):
    return {k: (map_tensors(v, fn) if isinstance(v, dict) else (fn(v) if torch.is_tensor(v) else v))
            for k, v in batch.items()}


def to_device(batch, device, non_blocking=True):
    return map_tensors(batch, lambda t: t.to(device, non_blocking=non_blocking))


def batch_nbytes(batch):
    n = 0
    for v in batch.values():
        if isinstance(v, dict):
            n += batch_nbytes(v)
        elif torch.is_tensor(v):
            n += v.numel() * v.element_size()
    return n
