"""oracle/gen_golden_dp.py -- TEST INFRASTRUCTURE.  Run in the BUILD CONTAINER only (needs
/root/reference; the GPU box never sees it):

    python -m oracle.gen_golden_dp            # writes tests/golden/dp_*.npz

Imports the REFERENCE's own Diffusion-Policy modules read-only from /root/reference --
`src/models/components/diffusion_policy/{diffusion_unet_image_policy,base_image_policy}.py`,
`.../diffusion/{conditional_unet1d,conv1d_components,positional_embedding,mask_generator}.py`,
`.../vision/pcd_obs_encoder.py`, `src/utils/diffusion_policy/*`, `src/utils/pytorch_utils.py` --
and runs `DiffusionUnetImagePolicy.compute_loss` forward + backward on small seeded inputs on the
CPU.  Substituted because the container lacks them (nothing else is):
  * everything gen_golden_act.py substitutes (pointops._C -> C oracle, legacy tensor constructors,
    `src.utils` package __init__, spconv PointNet -> `OraclePointNet`);
  * `src.utils.RankedLogger` (lightning) -> stdlib logger; `zarr` (import-only in normalizer.py) ->
    empty stub; `torchvision` import in pcd_obs_encoder.py is real (installed) or stubbed;
  * `diffusers.schedulers.scheduling_ddpm.DDPMScheduler` (diffusers 0.29.0 is not installed) ->
    `oracle.dp_oracle.DDPMSchedule`, the restatement of its published algorithm.  The scheduler's
    own parity is therefore UNPINNED; everything else in the fixture is the reference's arithmetic;
  * `multi_image_obs_encoder` (imports torchvision transforms / crop randomizer; only used as a type
    annotation by the policy) -> stub class.
The policy draws `noise = torch.randn(...)` and `timesteps = torch.randint(...)`
(diffusion_unet_image_policy.py:281-290); both calls are recorded and stored in the fixture so the
oracle / product replay exactly the same draw (`batch["_noise"]`, `batch["_timesteps"]`).
"""
from __future__ import annotations

import importlib
import logging
import sys
import types

import numpy as np
import torch

from .gen_golden_act import OUT, REF, _load, _pkg, install_reference_shim

CASES = {
    # scratch_pointnet_pcd.yaml structure at test size: PointNet(6 -> 32 classes) + SA(32 -> 32) + 1-layer projector,
    # U-Net down_dims (32, 64), kernel 5, 8 groups, FiLM scale+bias, horizon 16, 2 obs steps
    "maniskill_small": dict(qpos_dim=9, action_dim=7, backbone_classes=32, n_obs_steps=2, pcd_nsample=8, pcd_npoints=24,
                            pcd_hidden_dim=32, projector_layers=1, projector_channels=[32, 40, 40], horizon=16,
                            diffusion_step_embed_dim=32, down_dims=[32, 64], kernel_size=5, n_groups=8,
                            cond_predict_scale=True, goal_dim=0, batch=8, n=70),
    # three resolution levels (16 -> 8 -> 4, as the reference's [512, 1024, 2048]), language-goal embedding,
    # 2-layer projector
    "maniskill_goal": dict(qpos_dim=7, action_dim=7, backbone_classes=16, n_obs_steps=2, pcd_nsample=8, pcd_npoints=16,
                           pcd_hidden_dim=16, projector_layers=2, projector_channels=[16, 16, 32], horizon=16,
                           diffusion_step_embed_dim=32, down_dims=[32, 64, 128], kernel_size=5, n_groups=8,
                           cond_predict_scale=True, goal_dim=12, batch=8, n=50),
}


def install_dp_shim():
    act, tr, loss = install_reference_shim()
    utils = sys.modules["src.utils"]

    class RankedLogger(logging.LoggerAdapter):
        def __init__(self, name=__name__, rank_zero_only=False, extra=None):
            super().__init__(logging.getLogger(name), extra)

    utils.RankedLogger = RankedLogger
    if "zarr" not in sys.modules:
        try:
            importlib.import_module("zarr")
        except Exception:
            z = types.ModuleType("zarr")
            z.Array = type("Array", (), {})
            sys.modules["zarr"] = z
    try:
        importlib.import_module("torchvision")
    except Exception:
        sys.modules["torchvision"] = types.ModuleType("torchvision")
    _load("src.utils.pytorch_utils", REF / "src" / "utils" / "pytorch_utils.py")
    _pkg("src.utils.diffusion_policy", REF / "src" / "utils" / "diffusion_policy")
    for m in ("tensor_util", "dict_of_tensor_mixin", "module_attr_mixin", "normalizer", "shape_util"):
        try:
            mod = importlib.import_module("src.utils.diffusion_policy." + m)
        except Exception:
            if m in ("tensor_util", "shape_util"):
                continue
            raise
        for name in ("DictOfTensorMixin", "ModuleAttrMixin", "LinearNormalizer", "SingleFieldLinearNormalizer"):
            if hasattr(mod, name):
                setattr(sys.modules["src.utils.diffusion_policy"], name, getattr(mod, name))
    # diffusers stub -> the restated scheduler
    from .dp_oracle import DDPMSchedule

    d = _pkg("diffusers")
    ds = _pkg("diffusers.schedulers")
    dd = types.ModuleType("diffusers.schedulers.scheduling_ddpm")
    dd.DDPMScheduler = DDPMSchedule
    sys.modules["diffusers.schedulers.scheduling_ddpm"] = dd
    d.schedulers, ds.scheduling_ddpm = ds, dd
    base = REF / "src" / "models" / "components" / "diffusion_policy"
    _pkg("src.models.components.diffusion_policy", base)
    _pkg("src.models.components.diffusion_policy.diffusion", base / "diffusion")
    _pkg("src.models.components.diffusion_policy.vision", base / "vision")
    mi = types.ModuleType("src.models.components.diffusion_policy.vision.multi_image_obs_encoder")
    mi.MultiImageObsEncoder = type("MultiImageObsEncoder", (), {})
    sys.modules[mi.__name__] = mi
    pol = importlib.import_module("src.models.components.diffusion_policy.diffusion_unet_image_policy")
    enc = importlib.import_module("src.models.components.diffusion_policy.vision.pcd_obs_encoder")
    return pol, enc


def shape_meta_of(cfg):
    sm = {"obs": {"pcds": {"shape": [6], "type": "pcd"}, "qpos": {"shape": [cfg["qpos_dim"]], "type": "low_dim"}},
          "action": {"shape": [cfg["action_dim"]]}, "goal": None}
    if cfg["goal_dim"]:
        sm["goal"] = {"task_emb": {"shape": [cfg["goal_dim"]]}}
    return sm


def synth_dp_batch(cfg, seed):
    g = torch.Generator().manual_seed(seed)
    b, n_obs = cfg["batch"], cfg["n_obs_steps"]
    clouds = b * n_obs
    sizes = torch.randint(int(0.75 * cfg["n"]), cfg["n"] + 1, (clouds,), generator=g)
    total = int(sizes.sum())
    coord = torch.rand(total, 3, generator=g) - 0.5
    grid = torch.floor(coord / 0.005).long()
    grid = grid - grid.min(0).values
    color = torch.randint(0, 256, (total, 3), generator=g).float() / 127.5 - 1
    batch = {"obs": {"qpos": torch.randn(b, cfg["horizon"], cfg["qpos_dim"], generator=g),
                     "pcds": {"coord": coord, "grid_coord": grid, "feat": torch.cat([color, coord], 1),
                              "offset": torch.cumsum(sizes, 0)}},
             "action": torch.randn(b, cfg["horizon"], cfg["action_dim"], generator=g)}
    if cfg["goal_dim"]:
        batch["goal"] = {"task_emb": torch.randn(b, cfg["goal_dim"], generator=g)}
    return batch


def normalizer_fields(cfg, seed):
    """Non-trivial per-dimension affine normalisers (as `LinearNormalizer.fit` would leave them)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key, dim in (("qpos", cfg["qpos_dim"]), ("action", cfg["action_dim"])):
        out[key] = (0.5 + torch.rand(dim, generator=g), 0.2 * torch.randn(dim, generator=g))
    return out


def main():
    from .act_oracle import OraclePointNet
    from .dp_oracle import DDPMSchedule

    pol, enc = install_dp_shim()
    OUT.mkdir(parents=True, exist_ok=True)
    for name, cfg in CASES.items():
        torch.manual_seed(4242)
        sm = shape_meta_of(cfg)
        encoder = enc.PCDObsEncoder(shape_meta=sm, pcd_model=OraclePointNet(6, cfg["backbone_classes"]),
                                    share_pcd_model=True, n_obs_step=cfg["n_obs_steps"], pcd_nsample=cfg["pcd_nsample"],
                                    pcd_npoints=cfg["pcd_npoints"], use_mask=False, bg_ratio=0.0,
                                    pcd_hidden_dim=cfg["pcd_hidden_dim"], projector_layers=cfg["projector_layers"],
                                    projector_channels=cfg["projector_channels"])
        policy = pol.DiffusionUnetImagePolicy(shape_meta=sm, noise_scheduler=DDPMSchedule(num_train_timesteps=100),
                                              obs_encoder=encoder, horizon=cfg["horizon"], n_action_steps=8,
                                              n_obs_steps=cfg["n_obs_steps"], num_inference_steps=100,
                                              obs_as_global_cond=True,
                                              diffusion_step_embed_dim=cfg["diffusion_step_embed_dim"],
                                              down_dims=cfg["down_dims"], kernel_size=cfg["kernel_size"],
                                              n_groups=cfg["n_groups"], cond_predict_scale=cfg["cond_predict_scale"])
        import torch.nn as nn

        for key, (scale, offset) in normalizer_fields(cfg, 7).items():
            policy.normalizer.params_dict[key] = nn.ParameterDict(
                {"scale": nn.Parameter(scale, requires_grad=False), "offset": nn.Parameter(offset, requires_grad=False)})
        policy.train()
        # make every affine / bias parameter non-trivial so that its gradient path is exercised
        with torch.no_grad():
            for k, p in policy.named_parameters():
                if p.requires_grad and p.dim() == 1:
                    p.add_(0.1 * torch.randn_like(p))
        state = {k: v.detach().clone() for k, v in policy.state_dict().items()}
        batch = synth_dp_batch(cfg, 99)
        rec = {}
        real_randn, real_randint = torch.randn, torch.randint

        # ---- inference: the reference's predict_action (eval mode, 10 sampling steps) on the same observations,
        # BEFORE the training step touches the BatchNorm running statistics.  Every torch.randn draw of the
        # sampling loop (trajectory init, then one per step with t > 0) is recorded in order.
        draws = []

        def randn_rec(*a, **k):
            draws.append(real_randn(*a, **k))
            return draws[-1]

        policy.eval()
        policy.num_inference_steps = 10
        obs_in = {"obs": {"qpos": batch["obs"]["qpos"].clone(), "pcds": {k: v.clone() for k, v in batch["obs"]["pcds"].items()}}}
        if "goal" in batch:
            obs_in["goal"] = {"task_emb": batch["goal"]["task_emb"].clone()}
        torch.manual_seed(777)
        torch.randn = randn_rec
        try:
            with torch.no_grad():
                pred = policy.predict_action(obs_in)
        finally:
            torch.randn = real_randn
        policy.train()

        def randn(*a, **k):
            rec["noise"] = real_randn(*a, **k)
            return rec["noise"]

        def randint(*a, **k):
            rec["timesteps"] = real_randint(*a, **k)
            return rec["timesteps"]

        ref_batch = {"obs": {"qpos": batch["obs"]["qpos"].clone(), "pcds": {k: v.clone() for k, v in batch["obs"]["pcds"].items()}},
                     "action": batch["action"].clone()}
        if "goal" in batch:
            ref_batch["goal"] = {"task_emb": batch["goal"]["task_emb"].clone()}
        torch.manual_seed(31337)
        torch.randn, torch.randint = randn, randint
        try:
            out = policy.compute_loss(ref_batch)
        finally:
            torch.randn, torch.randint = real_randn, real_randint
        out["loss"].backward()
        flat = {"pred/action": pred["action"].numpy(), "pred/action_pred": pred["action_pred"].numpy(),
                "pred/noises": np.stack([d.numpy() for d in draws]),
                "meta/cfg_keys": np.array([k for k in cfg]), "meta/cfg_vals": np.array([repr(cfg[k]) for k in cfg]),
                "out/loss": out["loss"].detach().numpy(), "in/noise": rec["noise"].numpy(),
                "in/timesteps": rec["timesteps"].numpy()}
        from tests._golden_act import grad_summary

        nograd = []
        for k, p in policy.named_parameters():
            if not p.requires_grad:
                continue
            if p.grad is None:
                nograd.append(k)
            else:
                flat["grad/" + k] = grad_summary(p.grad)
        flat["meta/nograd"] = np.array(nograd, dtype=str)
        for k, v in state.items():
            flat["state/" + k] = v.numpy()
        for k, v in policy.state_dict().items():
            if "running_" in k:
                flat["post/" + k] = v.numpy()
        for k, v in batch["obs"]["pcds"].items():
            flat["in/obs/pcds/" + k] = v.numpy()
        flat["in/obs/qpos"], flat["in/action"] = batch["obs"]["qpos"].numpy(), batch["action"].numpy()
        if "goal" in batch:
            flat["in/goal/task_emb"] = batch["goal"]["task_emb"].numpy()
        path = OUT / f"dp_{name}.npz"
        np.savez_compressed(path, **flat)
        print(path, f"loss={float(out['loss']):.6f}", f"params={sum(p.numel() for p in policy.parameters())}",
              f"nograd={nograd}")


# Encoder-only fixtures for the set-abstraction variants of PCDObsEncoder (pcd_obs_encoder.py:91-93,133-177,
# 201-218): reference `PCDObsEncoder.forward` in train mode, loss = <output, fixed random tensor>.
ENC_CASES = {
    "mask": dict(qpos_dim=9, backbone_classes=32, pcd_nsample=8, pcd_npoints=24, pcd_hidden_dim=32, projector_layers=1,
                 projector_channels=[32, 40, 40], use_mask=True, bg_ratio=0.25, pre_sample=False, clouds=8, n=70),
    "presample": dict(qpos_dim=9, backbone_classes=32, pcd_nsample=8, pcd_npoints=24, pcd_hidden_dim=32, projector_layers=1,
                      projector_channels=[32, 40, 40], use_mask=False, bg_ratio=0.0, pre_sample=True, clouds=8, n=70),
}


def main_encoder():
    from .act_oracle import OraclePointNet
    from tests._golden_act import grad_summary

    pol, enc = install_dp_shim()
    for name, cfg in ENC_CASES.items():
        torch.manual_seed(99)
        sm = {"obs": {"pcds": {"shape": [6], "type": "pcd"}, "qpos": {"shape": [cfg["qpos_dim"]], "type": "low_dim"}}}
        e = enc.PCDObsEncoder(shape_meta=sm, pcd_model=OraclePointNet(6, cfg["backbone_classes"]), share_pcd_model=True,
                              n_obs_step=2, pcd_nsample=cfg["pcd_nsample"], pcd_npoints=cfg["pcd_npoints"],
                              use_mask=cfg["use_mask"], bg_ratio=cfg["bg_ratio"], pcd_hidden_dim=cfg["pcd_hidden_dim"],
                              projector_layers=cfg["projector_layers"], projector_channels=cfg["projector_channels"],
                              pre_sample=cfg["pre_sample"], in_channel=6).train()
        with torch.no_grad():
            for k, p_ in e.named_parameters():
                if p_.dim() == 1 and p_.numel():
                    p_.add_(0.1 * torch.randn_like(p_))
        state = {k: v.detach().clone() for k, v in e.state_dict().items()}
        g = torch.Generator().manual_seed(17)
        sizes = torch.randint(int(0.75 * cfg["n"]), cfg["n"] + 1, (cfg["clouds"],), generator=g)
        total = int(sizes.sum())
        coord = torch.rand(total, 3, generator=g) - 0.5
        grid = torch.floor(coord / 0.005).long()
        grid = grid - grid.min(0).values
        color = torch.randint(0, 256, (total, 3), generator=g).float() / 127.5 - 1
        pcds = {"coord": coord, "grid_coord": grid, "feat": torch.cat([color, coord], 1), "offset": torch.cumsum(sizes, 0),
                "mask": torch.rand(total, generator=g) < 0.55}
        qpos = torch.randn(cfg["clouds"], cfg["qpos_dim"], generator=g)
        out = e({"pcds": {k: v.clone() for k, v in pcds.items()}, "qpos": qpos})
        probe = torch.randn(out.shape, generator=g)
        (out * probe).sum().backward()
        flat = {"meta/cfg_keys": np.array(list(cfg)), "meta/cfg_vals": np.array([repr(cfg[k]) for k in cfg]),
                "out/features": out.detach().numpy(), "in/probe": probe.numpy(), "in/qpos": qpos.numpy()}
        for k, v in pcds.items():
            flat["in/pcds/" + k] = v.numpy()
        for k, v in state.items():
            flat["state/" + k] = v.numpy()
        for k, p_ in e.named_parameters():
            if p_.grad is not None:
                flat["grad/" + k] = grad_summary(p_.grad)
        for k, v in e.state_dict().items():
            if "running_" in k:
                flat["post/" + k] = v.numpy()
        path = OUT / f"dpenc_{name}.npz"
        np.savez_compressed(path, **flat)
        print(path, tuple(out.shape))


if __name__ == "__main__":
    import sys as _sys

    if "encoder" in _sys.argv[1:]:
        main_encoder()
    else:
        main()
