"""Host-side logic of the data-parallel training step (no GPU): OneCycle schedule vs torch,
flat parameter / gradient buffers, batch sharding, and the N>1 gradient exchange over gloo
(world_size 2, CPU) -- the collective is backend-agnostic: one SUM all-reduce of the flat buffer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from pointcloudmatters_b200.data import synthetic_act_batch
from pointcloudmatters_b200.trainer import BCTrainer, FlatState, OneCycle, shard_batch


def test_onecycle_matches_torch():
    p = [nn.Parameter(torch.zeros(1))]
    opt = torch.optim.AdamW(p, lr=5e-5, weight_decay=0.05)
    sch = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=5e-5, total_steps=300, pct_start=0.1, anneal_strategy="cos",
                                              div_factor=100.0, final_div_factor=1000.0)
    oc = OneCycle(5e-5, 300)
    for s in range(300):
        lr, b1 = oc.at(s)
        assert abs(lr - opt.param_groups[0]["lr"]) <= 1e-12
        assert abs(b1 - opt.param_groups[0]["betas"][0]) <= 1e-12
        opt.step()
        if s < 299:
            sch.step()


def test_flat_state_views_and_inactive_tail():
    torch.manual_seed(0)
    m = nn.Sequential(nn.Linear(5, 7), nn.Linear(7, 3), nn.Linear(3, 1))
    before = {k: v.clone() for k, v in m.state_dict().items()}
    m[0](torch.randn(2, 5)).sum().backward()  # only layer 0 receives a gradient
    inactive = [p for p in m.parameters() if p.grad is None]
    fs = FlatState(m.parameters(), inactive)
    assert fs.n_active == 40 + 8 and fs.param.numel() % 8 == 0  # (35->40) + (7->8), padded to 8
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    assert all(p.data.data_ptr() >= fs.param.data_ptr() for p in m.parameters())
    g0 = m[0].weight.grad.clone()
    fs.zero_grad()
    assert float(m[0].weight.grad.abs().sum()) == 0 and float(g0.abs().sum()) > 0
    m[0](torch.randn(2, 5)).sum().backward()  # autograd accumulates IN PLACE into the flat views
    assert float(fs.grad[: fs.n_active].abs().sum()) > 0
    assert m[0].weight.grad.data_ptr() == fs.grad.data_ptr()


def test_hyper_values_follow_adamw_bias_correction():
    tr = BCTrainer(nn.Linear(2, 2), lr=1e-3, total_steps=100)
    h0, h9 = tr.hyper_values(0), tr.hyper_values(9)
    assert abs(h0[5] - (1 - h0[1])) < 1e-12 and abs(h9[6] - (1 - 0.999 ** 10)) < 1e-12
    assert h0[7] == 0.5 and h0[8] == 1.0 and h0[4] == 0.05


def test_shard_batch_splits_by_sample():
    b = synthetic_act_batch(6, 50, num_queries=8, seed=3, ragged=True)
    shards = [shard_batch(b, r, 3) for r in range(3)]
    assert sum(s["qpos"].shape[0] for s in shards) == 6
    assert torch.equal(torch.cat([s["pcds"]["coord"] for s in shards]), b["pcds"]["coord"])
    assert torch.equal(torch.cat([s["actions"] for s in shards]), b["actions"])
    off = b["pcds"]["offset"]
    for r, s in enumerate(shards):
        assert s["pcds"]["offset"][-1] == s["pcds"]["coord"].shape[0]
        sizes = torch.diff(s["pcds"]["offset"], prepend=torch.zeros(1, dtype=off.dtype))
        full = torch.diff(off, prepend=torch.zeros(1, dtype=off.dtype))[2 * r: 2 * r + 2]
        assert torch.equal(sizes, full)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical replicas
    model = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 2))
    unused = nn.Linear(3, 3)  # never receives a gradient (like is_pad_head)
    params = list(model.parameters()) + list(unused.parameters())
    tr = BCTrainer(nn.ModuleList([model, unused]), lr=1e-3, total_steps=100)
    x = torch.randn(8, 4, generator=torch.Generator().manual_seed(7))
    xs = x[rank * 4:(rank + 1) * 4]  # this rank's shard of the global batch
    (model(xs).pow(2).sum() / 8).backward()
    tr._build_flat()
    tr.reduce_gradients()  # ONE collective over the flat buffer (SUM; 1/world folded into the optimizer)
    g = tr.flat.grad[: tr.flat.n_active].clone() / world
    q.put((rank, g, tr.flat.n_active, tr.flat.param.numel(), tr.hyper_values(0)[8]))
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    (r0, g0, na0, nt0, sc0), (r1, g1, na1, nt1, sc1) = res
    assert torch.equal(g0, g1) and na0 == na1 and sc0 == 0.5
    assert nt0 - na0 == 24  # the unused Linear(3,3): 9 -> 16 and 3 -> 8 padded floats parked in the inactive tail
    # equals the single-process gradient of the mean loss over the GLOBAL batch
    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(4, 8), nn.ReLU(), nn.Linear(8, 2))
    x = torch.randn(8, 4, generator=torch.Generator().manual_seed(7))
    (model(x).pow(2).sum() / 8).backward()
    ref = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 8)) for p in model.parameters()])
    torch.testing.assert_close(g0 * 2, ref, rtol=1e-5, atol=1e-6)  # each rank held half the samples of a /8 loss


def test_shard_batch_diffusion_policy_contract():
    """Data-parallel split of a Diffusion-Policy batch: n_obs_steps clouds per sample stay with their sample,
    offsets are rebased, per-shard `n_max` is exact, shards tile the global batch."""
    from pointcloudmatters_b200.data import synthetic_dp_batch
    from pointcloudmatters_b200.trainer import shard_batch

    g = synthetic_dp_batch(6, 50, n_obs_steps=2, goal_dim=4, seed=9, ragged=True)
    shards = [shard_batch(g, r, 3) for r in range(3)]
    assert all(s["action"].shape[0] == 2 and s["obs"]["qpos"].shape[0] == 2 and s["goal"]["task_emb"].shape[0] == 2 for s in shards)
    assert all(s["obs"]["pcds"]["offset"].shape[0] == 4 for s in shards)
    assert torch.equal(torch.cat([s["obs"]["pcds"]["coord"] for s in shards]), g["obs"]["pcds"]["coord"])
    assert torch.equal(torch.cat([s["action"] for s in shards]), g["action"])
    sizes = torch.diff(g["obs"]["pcds"]["offset"], prepend=torch.zeros(1, dtype=torch.int64))
    for r, s in enumerate(shards):
        want = sizes[4 * r: 4 * r + 4]
        assert torch.equal(torch.diff(s["obs"]["pcds"]["offset"], prepend=torch.zeros(1, dtype=torch.int64)), want)
        assert s["obs"]["pcds"]["n_max"] == int(want.max())
        assert int(s["obs"]["pcds"]["offset"][-1]) == s["obs"]["pcds"]["coord"].shape[0]


class _ToyPolicy(nn.Module):
    """Two-stage policy with a gradient-bucket boundary between the stages (like decoder | encoder | rest)."""

    def __init__(self):
        super().__init__()
        self.front, self.back = nn.Linear(4, 8), nn.Linear(8, 2)
        self.unused = nn.Linear(3, 3)

    def grad_buckets(self):
        return [("back", ("back.",)), (None, ("",))]

    def forward(self, batch):
        from pointcloudmatters_b200 import functional as PF

        h = torch.relu(self.front(batch["x"]))
        PF.grad_boundary(h, "back")  # d(h) available = self.back's gradients are final
        return {"loss": self.back(h).pow(2).sum() / 8}


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    pol = _ToyPolicy()
    tr = BCTrainer(pol, lr=1e-3, total_steps=100, input_keys=("x",), loss_keys=("loss",), overlap_allreduce=True)
    x = torch.randn(8, 4, generator=torch.Generator().manual_seed(7))
    batch = {"x": x[rank * 4:(rank + 1) * 4]}
    tr._forward_backward(batch)  # eager, no flat state yet: plain autograd
    tr._build_flat()
    launched = []
    orig = tr._launch_buckets
    tr._launch_buckets = lambda upto: (launched.append(upto), orig(upto))[1]
    tr._forward_backward(batch)  # zeroes the flat gradient, exchanges bucket by bucket during backward
    names = [t for t, _s, _e in tr.bucket_ranges]
    q.put((rank, tr.flat.grad[: tr.flat.n_active].clone(), names, list(tr.bucket_ranges), launched, tr._reduced,
           pol.back.weight.grad.data_ptr() == tr.flat.grad.data_ptr()))
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    (_, g0, names, ranges, launched, reduced, back_first), (_, g1, *_rest) = res
    assert torch.equal(g0, g1) and reduced
    assert names == ["back", None] and back_first  # the flat layout follows the bucket (completion) order
    assert ranges[0][2] - ranges[0][1] == 16 + 8 and ranges[1][1] == ranges[0][2]  # back: (2x8 -> 16) + (2 -> 8)
    assert launched[0] == 1 and launched[-1] == 2  # bucket 0 went out at its boundary, the rest after backward
    # equals the single-process gradient of the mean loss over the GLOBAL batch (sum over ranks of /8 losses)
    torch.manual_seed(0)
    pol = _ToyPolicy()
    x = torch.randn(8, 4, generator=torch.Generator().manual_seed(7))
    pol({"x": x})["loss"].backward()
    pad = lambda t: torch.nn.functional.pad(t.reshape(-1), (0, (-t.numel()) % 8))
    ref = torch.cat([pad(pol.back.weight.grad), pad(pol.back.bias.grad), pad(pol.front.weight.grad), pad(pol.front.bias.grad)])
    torch.testing.assert_close(g0, ref, rtol=1e-5, atol=1e-6)
