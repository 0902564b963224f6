"""`training_step` surface of the reference's LightningModules, without Lightning:
mirror of `ManiSkill2ACTBCModule` / `RLBenchACTBCModule`
(src/models/maniskill2_act_bc_module.py:16-86,347-367; src/models/rlbench_act_bc_module.py:60-82).

`training_step(batch, batch_idx) -> loss` and `configure_optimizers()` keep the reference's names
and meaning; validation rollouts (simulators) are out of scope.  Under Lightning the reference
module can be used unchanged with the policy swapped to `pointcloudmatters_b200.act.ACTPCD`
(INTEGRATION.md); this class is what the in-repo harness and bench.py drive directly, with the
DDP all-reduce / clip / AdamW / OneCycleLR that Lightning's Trainer would add folded into
`BCTrainer` (pointcloudmatters_b200/trainer.py).
"""
from __future__ import annotations

from typing import Any

import torch
import torch.nn as nn

from .trainer import BCTrainer


class ACTBCModule(nn.Module):
    def __init__(self, policy, optimizer: dict | None = None, lr_scheduler: dict | None = None,
                 gradient_clip_val: float = 0.5, total_steps: int = 100000, use_cuda_graph: bool = False,
                 accumulate_grad_batches: int = 1, debug_hints: bool = False, sync_batchnorm: bool = False,
                 overlap_allreduce: bool = True, grad_wire_dtype=None, **kwargs):
        super().__init__()
        self.policy = policy
        # defaults = configs/model/maniskill2_act_pcd_model.yaml:11-25, configs/trainer/ddp.yaml:12
        self.hparams = dict(optimizer=dict(type="AdamW", lr=5e-5, weight_decay=0.05) | (optimizer or {}),
                            lr_scheduler=lr_scheduler, gradient_clip_val=gradient_clip_val, total_steps=total_steps,
                            use_cuda_graph=use_cuda_graph, accumulate_grad_batches=accumulate_grad_batches,
                            debug_hints=debug_hints, sync_batchnorm=sync_batchnorm, overlap_allreduce=overlap_allreduce,
                            grad_wire_dtype=grad_wire_dtype)
        self._trainer: BCTrainer | None = None
        self.logged: dict[str, Any] = {}

    def forward(self, x):
        return self.policy(x)

    def model_step(self, batch):
        return self.policy(batch)

    def configure_optimizers(self) -> BCTrainer:
        opt = self.hparams["optimizer"]
        if opt.get("type", "AdamW") != "AdamW":
            raise NotImplementedError("the fused optimizer implements the reference's configured AdamW")
        sch = (self.hparams["lr_scheduler"] or {}).get("scheduler", {}) if self.hparams["lr_scheduler"] else {}
        sch = {k: v for k, v in sch.items() if k in ("pct_start", "div_factor", "final_div_factor")}
        self._trainer = BCTrainer(self.policy, lr=opt["lr"], weight_decay=opt.get("weight_decay", 0.01),
                                  clip_norm=self.hparams["gradient_clip_val"], total_steps=self.hparams["total_steps"],
                                  scheduler=sch, use_cuda_graph=self.hparams["use_cuda_graph"],
                                  accumulate_grad_batches=self.hparams["accumulate_grad_batches"],
                                  debug_hints=self.hparams["debug_hints"], **self._dist_kwargs())
        return self._trainer

    def _dist_kwargs(self):
        return {k: self.hparams[k] for k in ("sync_batchnorm", "overlap_allreduce", "grad_wire_dtype")}

    def prefetch(self, batch) -> None:
        """Stage the next step's pinned host batch on the device while the current step runs (BCTrainer.prefetch)."""
        if self._trainer is not None:
            self._trainer.prefetch(batch)

    def training_step(self, batch, batch_idx: int = 0) -> torch.Tensor:
        """forward + backward + gradient all-reduce + clip + AdamW + LR step; returns the loss."""
        if self._trainer is None:
            self.configure_optimizers()
        losses = self._trainer.training_step(batch)
        self.logged = {"train/loss": losses["loss"], "train/action_loss": losses["action_loss"],
                       "train/kl_loss": losses["kl_loss"]}
        return losses["loss"]


class DiffusionPolicyBCModule(ACTBCModule):
    """Mirror of `ManiSkill2DiffusionPolicyBCModule` (src/models/maniskill2_dp_bc_module.py:20-99):
    `model_step` = `policy.compute_loss(batch)` in training mode, `training_step` returns `loss_dict["loss"]`;
    optimizer / scheduler defaults = configs/model/maniskill2_diffusion_policy_model.yaml:10-25
    (AdamW lr 1e-4, betas (0.9, 0.95), weight decay 1e-4; OneCycleLR pct_start 0.15)."""

    def __init__(self, policy, optimizer: dict | None = None, lr_scheduler: dict | None = None,
                 gradient_clip_val: float = 0.5, total_steps: int = 100000, use_cuda_graph: bool = False,
                 accumulate_grad_batches: int = 1, debug_hints: bool = False, **kwargs):
        super().__init__(policy, dict(type="AdamW", lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.95)) | (optimizer or {}),
                         lr_scheduler or {"scheduler": dict(pct_start=0.15, div_factor=100.0, final_div_factor=1000.0)},
                         gradient_clip_val, total_steps, use_cuda_graph, accumulate_grad_batches, debug_hints, **kwargs)

    def setup(self, normalizer) -> None:
        """maniskill2_dp_bc_module.py:57-60: copy the dataset's fitted normaliser into the policy."""
        self.policy.set_normalizer(normalizer)

    def configure_optimizers(self) -> BCTrainer:
        opt = self.hparams["optimizer"]
        sch = {k: v for k, v in self.hparams["lr_scheduler"].get("scheduler", {}).items()
               if k in ("pct_start", "div_factor", "final_div_factor")}
        self._trainer = BCTrainer(self.policy, lr=opt["lr"], weight_decay=opt.get("weight_decay", 0.01),
                                  betas=tuple(opt.get("betas", (0.9, 0.999))), clip_norm=self.hparams["gradient_clip_val"],
                                  total_steps=self.hparams["total_steps"], scheduler=sch,
                                  use_cuda_graph=self.hparams["use_cuda_graph"],
                                  accumulate_grad_batches=self.hparams["accumulate_grad_batches"],
                                  debug_hints=self.hparams["debug_hints"], **self._dist_kwargs(),
                                  input_keys=("obs", "action", "goal", "_noise", "_timesteps"), loss_keys=("loss",))
        return self._trainer

    def training_step(self, batch, batch_idx: int = 0) -> torch.Tensor:
        if self._trainer is None:
            self.configure_optimizers()
        losses = self._trainer.training_step(batch)
        self.logged = {"train/loss": losses["loss"]}
        return losses["loss"]
