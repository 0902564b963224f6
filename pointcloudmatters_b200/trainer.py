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
from collections import OrderedDict
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
                 use_cuda_graph: bool = False, input_keys=None, loss_keys=None, accumulate_grad_batches: int = 1,
                 max_cached_graphs: int = 4, debug_hints: bool = False, sync_batchnorm: bool = False,
                 overlap_allreduce: bool = True, grad_wire_dtype=None):
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
        # shape signature -> (graph, static batch, static outputs); LRU-bounded: every entry owns a private activation
        # pool, and ragged real data produces a new signature (sum N of the packed cloud) for almost every batch
        self._graphs = OrderedDict()
        self.max_cached_graphs = max(1, int(max_cached_graphs))
        self._graph_lookups = self._graph_misses = 0
        self.graph_disabled_reason = None
        self._eager_steps = 0
        # reference presets use accumulate_grad_batches 2 (exp_maniskill2_act_policy/base.yaml:26-28): gradients of
        # k micro-batches are summed in the flat buffer, ONE all-reduce + optimizer step closes the group
        self.accumulate_grad_batches = max(1, int(accumulate_grad_batches))
        self._micro = 0
        self.debug_hints = bool(debug_hints)
        self._hook_handle = None
        # multi-rank options.  sync_batchnorm: exact parity with the reference DDP preset (configs/trainer/ddp.yaml:9) -- one
        # small collective per BatchNorm layer each way (functional.SYNC_BN); default off = local statistics.
        # overlap_allreduce: the flat gradient is laid out in the policy's bucket order (grad_buckets()) and each bucket's
        # all-reduce starts as soon as its gradients are final (functional.grad_boundary), overlapping the remaining
        # backward; off = ONE all-reduce after backward.  grad_wire_dtype=torch.bfloat16: all-reduce a bf16 copy (half the
        # NVLink bytes, like DDP's bf16 compression hook; not bit-compatible with the fp32 exchange).
        self.sync_batchnorm = bool(sync_batchnorm)
        self.overlap_allreduce = bool(overlap_allreduce)
        self.grad_wire_dtype = grad_wire_dtype
        self.bucket_ranges = []  # [(tag, start, end)] element ranges of the flat gradient, in completion order
        self._works, self._launched, self._reduced, self._exchange_now = [], 0, False, False
        # weight-gradient GEMMs into the flat gradient are queued during backward and run as grouped launches at the
        # bucket boundaries (functional.DW_QUEUE); PCM_NO_GROUPED_DW=1 restores one launch per product (A/B switch)
        import os

        self.group_weight_grads = not bool(int(os.environ.get("PCM_NO_GROUPED_DW", "0")))
        # the fused AdamW kernel leaves the gradient buffer zeroed, so the next step starts accumulating without a separate
        # zero-fill pass (p.grad reads as zeros after training_step; `last_grad_norm` keeps the pre-clip norm).
        # optimizer_zeroes_grad=False restores torch's behaviour (clipped gradients stay in p.grad).
        self.optimizer_zeroes_grad = True
        self._grads_clean = False
        if self.sync_batchnorm:
            from . import functional as PF

            PF.set_sync_batchnorm(True, process_group)

    # -- gradient exchange: ONE collective over the flat buffer -----------------------------------
    def _all_reduce(self, g, async_op=False):
        """SUM all-reduce of one slice of the flat gradient (optionally through a bf16 wire copy)."""
        if self.grad_wire_dtype is None or self.grad_wire_dtype == g.dtype:
            return dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg, async_op=async_op)
        wire = g.to(self.grad_wire_dtype)
        dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.pg)
        g.copy_(wire)
        return None

    def reduce_gradients(self):
        """SUM all-reduce of the flat gradient; the 1/world average is folded into the optimizer kernel.  No-op when the
        buckets were already exchanged during backward (overlap_allreduce)."""
        if self.world > 1 and not self._reduced:
            self._all_reduce(self.flat.grad[: self.flat.n_active])
        self._reduced = False

    # -- overlapped exchange: buckets start as soon as their gradients are final ---------------------
    def _launch_buckets(self, upto):
        """Start the all-reduce of buckets [launched, upto) on NCCL's stream (it waits for the calling stream's work so
        far; the caller keeps computing)."""
        while self._launched < upto:
            _tag, s0, e0 = self.bucket_ranges[self._launched]
            if e0 > s0:
                w = self._all_reduce(self.flat.grad[s0:e0], async_op=True)
                if w is not None:
                    self._works.append(w)
            self._launched += 1

    def _on_boundary(self, tag):
        from . import functional as PF

        PF.flush_dw_queue()  # the bucket's queued weight gradients must be in the flat buffer before it is exchanged
        if not self._exchange_now:
            return
        for i, (t, _s, _e) in enumerate(self.bucket_ranges):
            if t == tag:
                self._launch_buckets(i + 1)
                return

    def broadcast_buffers(self, average=True):
        """BatchNorm running statistics diverge across ranks when statistics are local (sync_batchnorm=False): average them
        (or take rank 0's) -- call before saving a checkpoint so that it does not depend on which rank saves."""
        if self.world <= 1:
            return
        for b in self.policy.buffers():
            if b.dtype.is_floating_point:
                if average:
                    dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.pg)
                    b.div_(self.world)
                else:
                    dist.broadcast(b, src=0, group=self.pg)

    def _bucketed_parameters(self):
        """(parameters in bucket order, [(tag, [params])]).  Order = the policy's `grad_buckets()` (completion order of the
        backward) so that every bucket is ONE contiguous slice of the flat gradient."""
        named = [(n, p) for n, p in self.policy.named_parameters() if p.requires_grad]
        spec = self.policy.grad_buckets() if hasattr(self.policy, "grad_buckets") else [(None, ("",))]
        buckets = [(tag, []) for tag, _ in spec]
        for n, p in named:
            for i, (_tag, prefixes) in enumerate(spec):
                if any(n.startswith(pre) for pre in prefixes):
                    buckets[i][1].append(p)
                    break
            else:
                buckets[-1][1].append(p)
        return [p for _t, ps in buckets for p in ps], buckets

    def _build_flat(self):
        inactive = [p for p in self.policy.parameters() if p.requires_grad and p.grad is None]
        ordered, buckets = self._bucketed_parameters()
        self.flat = FlatState(ordered, inactive)
        inactive_ids = {id(p) for p in inactive}
        pad8 = lambda n: (n + 7) // 8 * 8
        off, self.bucket_ranges = 0, []
        for tag, ps in buckets:
            size = sum(pad8(p.numel()) for p in ps if id(p) not in inactive_ids)
            self.bucket_ranges.append((tag, off, off + size))
            off += size
        assert off == self.flat.n_active
        self._finish_flat()

    def _finish_flat(self):
        import weakref

        dev = self.flat.param.device
        # `policy.load_state_dict` copies into the fp32 masters behind the optimizer's back: refresh the bf16 operand
        # copy the GEMMs read (a stale shadow would silently keep computing with the old weights)
        if self._hook_handle is None and hasattr(self.policy, "register_load_state_dict_post_hook"):
            me = weakref.ref(self)

            def _resync(_module, _incompatible):
                t = me()
                if t is not None and t.flat is not None:
                    t.flat.sync_shadow()

            self._hook_handle = self.policy.register_load_state_dict_post_hook(_resync)
        self._hyper_host = torch.zeros(9, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(9)
        self._hyper = torch.zeros(9, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def hyper_values(self, step_num: int):
        lr, beta1 = self.schedule.at(min(step_num, self.schedule.total_steps - 1))
        t = step_num + 1
        return [lr, beta1, self.betas[1], self.eps, self.weight_decay, 1.0 - beta1 ** t, 1.0 - self.betas[1] ** t,
                self.clip_norm if self.clip_norm else 0.0, 1.0 / (self.world * self.accumulate_grad_batches)]

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
                           self._sumsq, self._norm, f.param_bf16[: f.n_active], zero_grad=self.optimizer_zeroes_grad)
        self._grads_clean = self.optimizer_zeroes_grad  # the kernel left the flat gradient zeroed (the inactive tail never changes)
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

    def _forward_backward(self, batch, zero=True):
        from . import functional as PF

        if self.flat is not None and zero:
            self.flat.zero_grad()
        self._grads_clean = False
        with PF.stage("forward"):
            out = self.policy(self._inputs_only(batch))
        exchange = (self.world > 1 and self.overlap_allreduce and self.flat is not None and self.bucket_ranges
                    and self._micro + 1 >= self.accumulate_grad_batches and self.accumulate_grad_batches == 1)
        self._exchange_now = exchange
        if exchange:
            self._works, self._launched = [], 0
        group_dw = self.group_weight_grads and self.flat is not None and self.flat.param.is_cuda
        if group_dw:
            PF.DW_QUEUE = []
        if exchange or group_dw:
            PF.GRAD_BOUNDARY_CB = self._on_boundary
        try:
            with PF.stage("backward"):
                out["loss"].backward()
            if group_dw:
                PF.flush_dw_queue(final=True)
        finally:
            PF.GRAD_BOUNDARY_CB = None
            PF.DW_QUEUE = None
        if exchange:
            self._launch_buckets(len(self.bucket_ranges))
            for w in self._works:
                w.wait()
            self._works = []
            self._reduced = True
        return {k: out[k].detach() for k in self.LOSS_KEYS}

    HINT_KEYS = ("n_max", "fg_n_max", "bg_n_max")
    HINT_BUCKET = 128

    @classmethod
    def _bucket_hints(cls, batch):
        """Cloud-size hints are UPPER bounds for the kernels (FPS sizes its per-thread strips from them and reads
        only rows below the offsets): rounding them up to a multiple of HINT_BUCKET keeps batches whose largest
        cloud differs by a few points on ONE captured graph."""
        out = {}
        for k, v in batch.items():
            if isinstance(v, dict):
                out[k] = cls._bucket_hints(v)
            elif k in cls.HINT_KEYS and isinstance(v, int):
                out[k] = -(-v // cls.HINT_BUCKET) * cls.HINT_BUCKET
            else:
                out[k] = v
        return out

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

    def _graphed_forward_backward(self, batch, zero=True):
        """Replay (capturing on first use per batch-shape signature) the forward+backward graph.
        ~900 kernel launches per step become one cudaGraphLaunch; inputs are copied into static
        buffers, dropout seeds come from device memory so every replay draws fresh masks.
        The cache holds at most `max_cached_graphs` entries (least recently used evicted -- its private memory pool
        is released with it); when captures keep missing (ragged clouds: a new sum-N almost every batch) the
        trainer stops capturing and runs eagerly (`graph_disabled_reason` says so) instead of re-capturing each step."""
        batch_obj = batch
        batch = self._bucket_hints(self._inputs_only(batch))
        sig = (self._signature(batch), bool(zero))
        self._graph_lookups += 1
        entry = self._graphs.get(sig)
        if entry is None:
            self._graph_misses += 1
            if self._graph_lookups >= 16 and self._graph_misses > 0.5 * self._graph_lookups:
                self.graph_disabled_reason = (f"{self._graph_misses} captures in {self._graph_lookups} steps: batch shapes "
                                              "do not repeat (ragged clouds); running eagerly")
                self._graphs.clear()
                return self._forward_backward(self._on_device(batch), zero)
            while len(self._graphs) >= self.max_cached_graphs:
                self._graphs.popitem(last=False)
            dev = self.flat.param.device  # (a pinned host batch is staged straight into the static inputs, see training_step)
            clone = lambda v: ({kk: clone(vv) for kk, vv in v.items()} if isinstance(v, dict)
                               else (v.to(dev, copy=True) if torch.is_tensor(v) else v))
            static = clone(batch)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self._forward_backward(static, zero)
            entry = (graph, static, outs, self._reduced)  # was the (bucketed) gradient exchange captured with the step?
            self._graphs[sig] = entry
        else:
            self._graphs.move_to_end(sig)
        graph, static, outs, reduced_inside = entry
        pf, self._prefetched = getattr(self, "_prefetched", None), None
        if pf is not None and pf[0] is batch_obj and pf[1] == sig[0]:
            # the batch was staged host->device by `prefetch` while the previous step ran: device->device into the graph inputs
            main = torch.cuda.current_stream()
            main.wait_event(pf[3])
            self._copy_into(static, pf[2])
            done = torch.cuda.Event()
            done.record(main)
            self._stage_free[pf[1]] = done
        else:
            self._copy_into(static, batch)
        graph.replay()
        self._reduced = reduced_inside
        return outs

    def prefetch(self, batch):
        """Optional input pipelining for the CUDA-graph path: stage the NEXT step's pinned host batch host->device on a
        copy stream while the current step is still running.  `training_step(batch)` called afterwards with the SAME
        batch object then only copies device->device into the graph's inputs.  A no-op for device batches, on the eager
        path and before the step graphs exist; a prefetched batch that is never used is simply dropped."""
        if not (self.use_cuda_graph and self.flat is not None and self.flat.param.is_cuda and self.graph_disabled_reason is None):
            return
        b = self._bucket_hints(self._inputs_only(batch))
        host = []
        self._walk(b, lambda t: host.append(t.device.type == "cpu"))
        if not host or not any(host):
            return
        key = self._signature(b)
        if not hasattr(self, "_stages"):
            self._stages, self._stage_free, self._copy_stream = {}, {}, torch.cuda.Stream()
        dev = self.flat.param.device
        stage = self._stages.get(key)
        if stage is None:
            if len(self._stages) >= 4:
                self._stages.clear()
                self._stage_free.clear()
            mk = lambda v: ({kk: mk(vv) for kk, vv in v.items()} if isinstance(v, dict)
                            else (torch.empty(v.shape, dtype=v.dtype, device=dev) if torch.is_tensor(v) else v))
            stage = self._stages[key] = mk(b)
            # once: the staging buffers were just allocated on the step's stream (they are kept for the trainer's lifetime)
            self._copy_stream.wait_stream(torch.cuda.current_stream())
        cs = self._copy_stream
        free = self._stage_free.get(key)
        if free is not None:
            cs.wait_event(free)  # the previous step's device->device copy out of this staging set has finished
        with torch.cuda.stream(cs):
            self._copy_into(stage, b)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._prefetched = (batch, key, stage, ev)

    @staticmethod
    def _walk(batch, fn):
        for v in batch.values():
            if isinstance(v, dict):
                BCTrainer._walk(v, fn)
            elif torch.is_tensor(v):
                fn(v)

    def _on_device(self, batch):
        """Host (ideally pinned) batch -> this rank's device, asynchronously; device batches pass through."""
        dev = self.flat.param.device if self.flat is not None else next(self.policy.parameters()).device
        mv = lambda v: ({kk: mv(vv) for kk, vv in v.items()} if isinstance(v, dict)
                        else (v.to(dev, non_blocking=True) if torch.is_tensor(v) and v.device != dev else v))
        return mv(batch)

    def _check_hints(self, pcds):
        """debug_hints=True: one device->host read per step that verifies the host-provided cloud-size hints really are
        upper bounds (fps.cu clamps the cloud to the hint, so an undersized `n_max` would silently drop points)."""
        off = pcds["offset"]
        sizes = torch.diff(off, prepend=off.new_zeros(1))
        if pcds.get("n_max", None) is not None and int(sizes.max()) > int(pcds["n_max"]):
            raise ValueError(f"pcds['n_max']={pcds['n_max']} is smaller than the largest cloud ({int(sizes.max())})")

    def training_step(self, batch):
        """One full step on this rank's shard; returns the (detached) loss dict.  With `accumulate_grad_batches`
        = k the first k-1 calls of a group only add their gradients into the flat buffer; the k-th also runs the
        all-reduce, clip and AdamW (the kernel divides the summed gradient by world * k).
        `batch` may live in (pinned) host memory: on the graph path its tensors are copied host->device straight into
        the captured graph's static inputs (one asynchronous copy per tensor, no intermediate device batch); on the
        eager path it is moved to the device first."""
        from . import functional as PF

        PF.DROPOUT_RNG.new_step()
        # graph capture needs a step without device->host reads: the policy says whether this batch
        # carries the host-known cloud-size hints that make FPS sync-free (act.ACTPCD.sync_free)
        sync_free = getattr(self.policy, "sync_free", None)
        pcds = batch.get("pcds", None) if "pcds" in batch else batch.get("obs", {}).get("pcds", None)
        if self.debug_hints and pcds is not None:
            self._check_hints(pcds)
        graph_ok = sync_free is None or pcds is None or sync_free(pcds)
        zero = self._micro == 0 and not self._grads_clean
        if (self.use_cuda_graph and graph_ok and self.flat is not None and self._eager_steps >= 2
                and self.graph_disabled_reason is None):
            losses = self._graphed_forward_backward(batch, zero)
        else:
            if self.flat is None and not zero:
                raise RuntimeError("internal: flat state must exist before an accumulation group continues")
            losses = self._forward_backward(self._on_device(batch), zero)
            self._eager_steps += 1
            if self.flat is None:  # first step: discover never-used parameters, then go flat
                self._build_flat()
        self._grads_clean = False  # (a graph replay does not run _forward_backward's bookkeeping)
        self._micro += 1
        if self._micro >= self.accumulate_grad_batches:
            self._micro = 0
            with PF.stage("all-reduce (exposed part)"):
                self.reduce_gradients()
            with PF.stage("clip + AdamW"):
                self.optimizer_step()
        return losses

    # -- checkpoint surface (Lightning saves optimizer + scheduler state next to the weights) ---------------
    def state_dict(self):
        """Optimizer / schedule state: Adam moments per parameter NAME (layout-independent), the step counter and the
        dropout seed base.  Weights travel in `policy.state_dict()` as usual."""
        from . import functional as PF

        if self.flat is None:
            return {"step_num": self.step_num, "exp_avg": {}, "exp_avg_sq": {}, "dropout_seed": None}
        names = {id(p): n for n, p in self.policy.named_parameters()}
        f = self.flat
        avg, sq = {}, {}
        off = 0
        for p in f.active:
            n = p.numel()
            avg[names[id(p)]] = f.exp_avg[off:off + n].view_as(p).detach().clone()
            sq[names[id(p)]] = f.exp_avg_sq[off:off + n].view_as(p).detach().clone()
            off += (n + 7) // 8 * 8
        seed = PF.DROPOUT_RNG.base.get(str(f.param.device))
        return {"step_num": self.step_num, "exp_avg": avg, "exp_avg_sq": sq,
                "dropout_seed": None if seed is None else seed.detach().cpu().clone(),
                "inactive": [names[id(p)] for p in f.inactive]}

    def load_state_dict(self, state):
        """Restore what `state_dict` saved (call after `policy.load_state_dict`).  Builds the flat buffers if the trainer
        has not stepped yet; the bf16 operand copy is refreshed from the fp32 masters."""
        from . import functional as PF

        if self.flat is None:
            named = dict(self.policy.named_parameters())
            inactive = [named[n] for n in state.get("inactive", []) if n in named]
            self.flat = FlatState(self.policy.parameters(), inactive)
            self._finish_flat()
            self._eager_steps = max(self._eager_steps, 2)
        names = {id(p): n for n, p in self.policy.named_parameters()}
        f = self.flat
        off = 0
        for p in f.active:
            n = p.numel()
            k = names[id(p)]
            if k in state["exp_avg"]:
                f.exp_avg[off:off + n].copy_(state["exp_avg"][k].reshape(-1))
                f.exp_avg_sq[off:off + n].copy_(state["exp_avg_sq"][k].reshape(-1))
            off += (n + 7) // 8 * 8
        self.step_num = int(state["step_num"])
        if state.get("dropout_seed", None) is not None and f.param.is_cuda:
            PF.DROPOUT_RNG.base_for(f.param.device).copy_(state["dropout_seed"])
        f.sync_shadow()
        self._graphs.clear()


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
