"""Data-parallel BC training step -- the B200 replacement for what Lightning + DDP + torch.optim do
around the reference's `training_step` (configs/trainer/ddp.yaml:4-15,
src/models/maniskill2_act_bc_module.py:69-86,347-367):

    forward -> backward -> ONE all-reduce of a flat fp32 gradient buffer (NCCL over NVLink)
            -> clip-by-global-norm(0.5) -> AdamW (lr 5e-5, wd 0.05) -> OneCycleLR step

All parameters are views into one flat fp32 buffer and all gradients views into another, so the
gradient exchange is a single collective and the optimizer is a single fused kernel launch over
contiguous memory (no per-parameter loops, no host sync: the clip coefficient stays on device).
State parity with the reference optimizer: parameters that never receive a gradient
(`is_pad_head.*`) are skipped entirely like torch.optim.AdamW skips `grad is None`;
the dead decoder layers receive exact zeros and are therefore only weight-decayed.
"""
from __future__ import annotations

import math
from typing import Iterable

import torch
import torch.distributed as dist
import torch.nn as nn


class OneCycle:
    """torch.optim.lr_scheduler.OneCycleLR (two-phase, cos) as a pure function of the step index;
    config of configs/model/maniskill2_act_pcd_model.yaml:16-25 (reference wrapper
    src/utils/scheduler.py:102-139 keeps torch's defaults cycle_momentum=True, 0.85 / 0.95)."""

    def __init__(self, max_lr, total_steps, pct_start=0.1, div_factor=100.0, final_div_factor=1000.0,
                 base_momentum=0.85, max_momentum=0.95, anneal_strategy="cos"):
        assert anneal_strategy == "cos"
        self.total_steps = int(total_steps)
        initial = max_lr / div_factor
        self.phases = [
            (float(pct_start * total_steps) - 1, initial, max_lr, max_momentum, base_momentum),
            (total_steps - 1, max_lr, initial / final_div_factor, base_momentum, max_momentum),
        ]

    @staticmethod
    def _cos(start, end, pct):
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1)

    def at(self, step_num: int):
        """(lr, beta1) in force for optimizer step number `step_num` (0-based)."""
        start_step = 0.0
        for i, (end_step, lr0, lr1, m0, m1) in enumerate(self.phases):
            if step_num <= end_step or i == len(self.phases) - 1:
                pct = (step_num - start_step) / (end_step - start_step)
                return self._cos(lr0, lr1, pct), self._cos(m0, m1, pct)
            start_step = end_step
        raise AssertionError


class FlatState:
    """Flat fp32 parameter / gradient / Adam-moment buffers with per-parameter views."""

    def __init__(self, params: Iterable[nn.Parameter], inactive: Iterable[nn.Parameter] = ()):
        inactive_ids = {id(p) for p in inactive}
        params = [p for p in params if p.requires_grad]
        self.active = [p for p in params if id(p) not in inactive_ids]
        self.inactive = [p for p in params if id(p) in inactive_ids]
        ordered = self.active + self.inactive
        dev = ordered[0].device
        # every view starts on a multiple of 8 elements: 16-byte aligned in the bf16 operand copy too
        # (TMA base alignment), 32-byte aligned in fp32
        pad8 = lambda n: (n + 7) // 8 * 8
        self.n_active = sum(pad8(p.numel()) for p in self.active)
        total = self.n_active + sum(pad8(p.numel()) for p in self.inactive)
        self.param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(self.n_active, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n_active, dtype=torch.float32, device=dev)
        # bf16 operand copy of the parameters (kept current by the fused AdamW kernel)
        self.param_bf16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        off = 0
        for p in ordered:
            n = p.numel()
            self.param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.param[off:off + n].view_as(p)
            old = p.grad
            p.grad = self.grad[off:off + n].view_as(p)
            if old is not None:
                p.grad.copy_(old)
            off += pad8(n)
        self.sync_shadow()

    def sync_shadow(self):
        """Refresh the bf16 operand copy from the fp32 masters (after construction or after weights
        were changed behind the optimizer's back, e.g. load_state_dict) and publish it to the ops."""
        from . import functional as PF

        self.param_bf16.copy_(self.param)
        if self.param.is_cuda:
            PF.BF16_SHADOW.register(self)

    def zero_grad(self):
        self.grad.zero_()


class BCTrainer:
    """forward + backward (+ CUDA graph) -> one all-reduce -> fused clip + AdamW -> LR step."""

    def __init__(self, policy: nn.Module, lr=5e-5, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8,
                 clip_norm=0.5, total_steps=100000, scheduler: dict | None = None, process_group=None,
                 use_cuda_graph: bool = False, input_keys=None, loss_keys=None):
        self.policy = policy
        if input_keys is not None:
            self.INPUT_KEYS = tuple(input_keys)
        if loss_keys is not None:
            self.LOSS_KEYS = tuple(loss_keys)
        self.lr, self.weight_decay, self.betas, self.eps, self.clip_norm = lr, weight_decay, betas, eps, clip_norm
        sch = dict(pct_start=0.1, div_factor=100.0, final_div_factor=1000.0)
        sch.update(scheduler or {})
        self.schedule = OneCycle(max_lr=lr, total_steps=total_steps, **sch)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.flat: FlatState | None = None
        self.step_num = 0
        self.last_grad_norm = None
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}  # shape signature -> (graph, static batch, static outputs)
        self._eager_steps = 0

    # -- gradient exchange: ONE collective over the flat buffer -----------------------------------
    def reduce_gradients(self):
        """SUM all-reduce of the flat gradient; the 1/world average is folded into the optimizer kernel."""
        if self.world > 1:
            dist.all_reduce(self.flat.grad[: self.flat.n_active], op=dist.ReduceOp.SUM, group=self.pg)

    def _build_flat(self):
        inactive = [p for p in self.policy.parameters() if p.requires_grad and p.grad is None]
        self.flat = FlatState(self.policy.parameters(), inactive)
        dev = self.flat.param.device
        self._hyper_host = torch.zeros(9, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(9)
        self._hyper = torch.zeros(9, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def hyper_values(self, step_num: int):
        lr, beta1 = self.schedule.at(min(step_num, self.schedule.total_steps - 1))
        t = step_num + 1
        return [lr, beta1, self.betas[1], self.eps, self.weight_decay, 1.0 - beta1 ** t, 1.0 - self.betas[1] ** t,
                self.clip_norm if self.clip_norm else 0.0, 1.0 / self.world]

    def optimizer_step(self):
        from . import functional as PF

        f = self.flat
        # a FRESH pinned staging vector per step: the caching host allocator does not hand the block out again
        # before the asynchronous copy that used it has completed, so a host running several (graph-replayed)
        # steps ahead of the device can never overwrite values a queued copy has yet to read
        vals = torch.tensor(self.hyper_values(self.step_num), dtype=torch.float32)
        self._hyper_host = vals.pin_memory() if f.param.is_cuda else vals
        self._hyper.copy_(self._hyper_host, non_blocking=True)
        PF.clip_adamw_step(f.param[: f.n_active], f.grad[: f.n_active], f.exp_avg, f.exp_avg_sq, self._hyper,
                           self._sumsq, self._norm, f.param_bf16[: f.n_active])
        self.last_grad_norm = self._norm
        self.step_num += 1

    # -- forward + backward -----------------------------------------------------------------------
    INPUT_KEYS = ("pcds", "qpos", "actions", "is_pad", "goal_cond", "env_state", "_eps")  # ACT batch contract
    LOSS_KEYS = ("loss", "action_loss", "kl_loss")

    @staticmethod
    def _copy_dicts(v):
        return {k: BCTrainer._copy_dicts(x) for k, x in v.items()} if isinstance(v, dict) else v

    def _inputs_only(self, batch):
        """The policy's forward writes its intermediates into the dict it is given (reference
        behaviour, act.py:137-309); work on a copy of the (nested) dicts restricted to the batch-contract keys."""
        return {k: self._copy_dicts(batch[k]) for k in self.INPUT_KEYS if k in batch}

    def _forward_backward(self, batch):
        if self.flat is not None:
            self.flat.zero_grad()
        out = self.policy(self._inputs_only(batch))
        out["loss"].backward()
        return {k: out[k].detach() for k in self.LOSS_KEYS}

    @staticmethod
    def _signature(batch):
        sig = []
        for k in sorted(batch):
            v = batch[k]
            if isinstance(v, dict):
                sig.append((k, BCTrainer._signature(v)))
            elif torch.is_tensor(v):
                sig.append((k, tuple(v.shape), str(v.dtype)))
            else:
                sig.append((k, v))
        return tuple(sig)

    @staticmethod
    def _copy_into(static, batch):
        for k, v in batch.items():
            if isinstance(v, dict):
                BCTrainer._copy_into(static[k], v)
            elif torch.is_tensor(v):
                static[k].copy_(v, non_blocking=True)

    def _graphed_forward_backward(self, batch):
        """Replay (capturing on first use per batch-shape signature) the forward+backward graph.
        ~2000 kernel launches per step become one cudaGraphLaunch; inputs are copied into static
        buffers, dropout seeds come from device memory so every replay draws fresh masks."""
        batch = self._inputs_only(batch)
        sig = self._signature(batch)
        entry = self._graphs.get(sig)
        if entry is None:
            clone = lambda v: ({kk: clone(vv) for kk, vv in v.items()} if isinstance(v, dict)
                               else (v.clone() if torch.is_tensor(v) else v))
            static = clone(batch)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self._forward_backward(static)
            entry = (graph, static, outs)
            self._graphs[sig] = entry
        graph, static, outs = entry
        self._copy_into(static, batch)
        graph.replay()
        return outs

    def training_step(self, batch):
        """One full step on this rank's shard; returns the (detached) loss dict."""
        from . import functional as PF

        PF.DROPOUT_RNG.new_step()
        # graph capture needs a step without device->host reads: the policy says whether this batch
        # carries the host-known cloud-size hints that make FPS sync-free (act.ACTPCD.sync_free)
        sync_free = getattr(self.policy, "sync_free", None)
        pcds = batch.get("pcds", None) if "pcds" in batch else batch.get("obs", {}).get("pcds", None)
        graph_ok = sync_free is None or pcds is None or sync_free(pcds)
        if self.use_cuda_graph and graph_ok and self.flat is not None and self._eager_steps >= 2:
            losses = self._graphed_forward_backward(batch)
        else:
            losses = self._forward_backward(batch)
            self._eager_steps += 1
            if self.flat is None:  # first step: discover never-used parameters, then go flat
                self._build_flat()
        self.reduce_gradients()
        self.optimizer_step()
        return losses


def _shard_clouds(v: dict, lo: int, hi: int) -> dict:
    """Clouds [lo, hi) of a packed point-cloud dict (cumulative `offset`), offsets rebased to the shard."""
    off = v["offset"]
    start = int(off[lo - 1]) if lo > 0 else 0
    end = int(off[hi - 1])
    pc = {kk: vv[start:end] for kk, vv in v.items() if torch.is_tensor(vv) and kk != "offset"}
    pc["offset"] = off[lo:hi] - start
    sizes = torch.diff(off[lo:hi], prepend=off.new_tensor([start]))
    for kk, vv in v.items():
        if not torch.is_tensor(vv):
            pc[kk] = int(sizes.max()) if kk == "n_max" else vv  # keep the sync-free hint exact for the shard
    return pc


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """DistributedSampler-style split of one collated global batch by SAMPLE (contiguous blocks):
    clouds are independent units, so no data-path collective is needed.
    Handles the ACT contract (`pcds` at the top level, one cloud per sample) and the Diffusion-Policy contract
    (`obs.pcds` with n_obs_steps clouds per sample, `action`, optional `goal`)."""
    key = "qpos" if "qpos" in batch else "action"
    b = batch[key].shape[0]
    assert b % world == 0, "global batch must divide evenly across ranks"
    per = b // world
    lo, hi = rank * per, (rank + 1) * per

    def split(v, clouds_per_sample=1):
        if isinstance(v, dict):
            if "offset" in v and "coord" in v:
                c = v["offset"].shape[0] // b
                return _shard_clouds(v, lo * c, hi * c)
            return {kk: split(vv) for kk, vv in v.items()}
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == b:
            return v[lo:hi]
        return v

    return {k: split(v) for k, v in batch.items()}
