"""oracle/gen_golden_act.py -- TEST INFRASTRUCTURE.  Run in the BUILD CONTAINER only
(needs /root/reference; the GPU box never sees it):

    python -m oracle.gen_golden_act [case ...]   # writes tests/golden/act_*.npz (all cases, or the named ones)

Imports the REFERENCE's own modules read-only from /root/reference --
`src/models/components/act/{act,transformer,utils}.py`, `src/models/components/loss/misc.py`,
`src/utils/{sparse_tensor_utils,rotation_conversions}.py` and the Python wrappers of
`libs/pointops/functions/` -- and runs `ACTPCD` / `ACTRLBenchPCD` forward + backward on small
seeded inputs on the CPU.  Three things the container lacks are substituted, nothing else:
  * `pointops._C` (CUDA extension): a stub module whose functions are the C oracle
    (oracle/pointops_oracle.c), itself pinned bit-exactly to the real kernels
    (tests/golden/ref_pointops_*.npz).  The reference's Python wrappers run unmodified on top;
  * `torch.cuda.IntTensor/FloatTensor` (legacy constructors the wrappers allocate outputs with,
    e.g. functions/sampling.py:17-18): mapped to their CPU equivalents for the duration;
  * `src.utils` package __init__ (imports lightning, hydra, ...): replaced by a stub exposing
    only `offset2batch` (loaded from the reference's own sparse_tensor_utils.py);
  * the spconv-based PointNet backbone (spconv is not installed): `OraclePointNet`, the
    Linear-chain restatement.
The only randomness in a dropout-free forward is `reparametrize`'s normal_() draw
(act/utils.py:36-39); it is reproduced by seeding torch's global generator right before forward.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(name, path=None):
    m = types.ModuleType(name)
    m.__path__ = [str(path)] if path else []
    sys.modules[name] = m
    return m


def install_reference_shim():
    """Make `import pointops` and `import src.models.components.act.act` resolve to the reference."""
    from . import pointops_oracle as PO

    L = PO.lib()

    def ptr(t):
        import ctypes

        return ctypes.c_void_p(t.data_ptr())

    C = types.ModuleType("pointops._C")

    def farthest_point_sampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
        L.oracle_farthest_point_sampling(int(b), int(n), ptr(xyz), ptr(offset), ptr(new_offset), ptr(tmp), ptr(idx))

    def knn_query_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
        L.oracle_knn_query(int(m), int(nsample), ptr(xyz), ptr(new_xyz), ptr(offset), ptr(new_offset), ptr(idx), ptr(dist2))

    C.farthest_point_sampling_cuda = farthest_point_sampling_cuda
    C.knn_query_cuda = knn_query_cuda
    for n in ["ball_query_cuda", "random_ball_query_cuda", "grouping_forward_cuda", "grouping_backward_cuda",
              "interpolation_forward_cuda", "interpolation_backward_cuda", "subtraction_forward_cuda",
              "subtraction_backward_cuda", "aggregation_forward_cuda", "aggregation_backward_cuda",
              "attention_relation_step_forward_cuda", "attention_relation_step_backward_cuda",
              "attention_fusion_step_forward_cuda", "attention_fusion_step_backward_cuda"]:
        setattr(C, n, None)  # not on the ACT path
    sys.modules["pointops._C"] = C
    torch.cuda.IntTensor = torch.IntTensor  # legacy constructors -> CPU
    torch.cuda.FloatTensor = torch.FloatTensor
    fdir = REF / "libs" / "pointops" / "functions"
    spec = importlib.util.spec_from_file_location("pointops", fdir / "__init__.py", submodule_search_locations=[str(fdir)])
    pointops = importlib.util.module_from_spec(spec)
    sys.modules["pointops"] = pointops
    spec.loader.exec_module(pointops)

    _pkg("src", REF / "src")
    utils = _pkg("src.utils", REF / "src" / "utils")
    stu = _load("src.utils.sparse_tensor_utils", REF / "src" / "utils" / "sparse_tensor_utils.py")
    utils.offset2batch = stu.offset2batch
    _load("src.utils.rotation_conversions", REF / "src" / "utils" / "rotation_conversions.py")
    _pkg("src.models", REF / "src" / "models")
    _pkg("src.models.components", REF / "src" / "models" / "components")
    _pkg("src.models.components.act", REF / "src" / "models" / "components" / "act")
    _pkg("src.models.components.loss", REF / "src" / "models" / "components" / "loss")
    act = importlib.import_module("src.models.components.act.act")
    tr = importlib.import_module("src.models.components.act.transformer")
    loss = importlib.import_module("src.models.components.loss.misc")
    return act, tr, loss


CASES = {
    # name: (rlbench, cfg, batch, n_points)
    "maniskill_small": (False, dict(hidden_dim=96, nhead=2, dim_feedforward=32, enc_layers=2, dec_layers=3,
                                    dropout=0.0, num_queries=10, action_dim=7, qpos_dim=9, goal_cond_dim=3,
                                    latent_dim=32, kl_weight=10.0, pcd_npoints=32, pcd_nsample=8), 3, 96),
    "rlbench_small": (True, dict(hidden_dim=48, nhead=2, dim_feedforward=32, enc_layers=1, dec_layers=2,
                                 dropout=0.0, num_queries=6, action_dim=11, qpos_dim=4, goal_cond_dim=16,
                                 latent_dim=32, kl_weight=10.0, pcd_npoints=24, pcd_nsample=16, collision=True,
                                 position_loss_weight=3.0), 2, 70),
    # SURVEY.md 8 a4': set-abstraction variants selected by config (act.py:366-376,396-442,509-527)
    "maniskill_presample": (False, dict(hidden_dim=96, nhead=2, dim_feedforward=32, enc_layers=1, dec_layers=2,
                                        dropout=0.0, num_queries=8, action_dim=7, qpos_dim=9, goal_cond_dim=3,
                                        latent_dim=32, kl_weight=10.0, pcd_npoints=32, pcd_nsample=8, pre_sample=1,
                                        backbone_classes=96), 3, 96),  # backbone must emit hidden_dim channels
    "maniskill_mask": (False, dict(hidden_dim=96, nhead=2, dim_feedforward=32, enc_layers=1, dec_layers=2,
                                   dropout=0.0, num_queries=8, action_dim=7, qpos_dim=9, goal_cond_dim=3,
                                   latent_dim=32, kl_weight=10.0, pcd_npoints=32, pcd_nsample=8, use_mask=1,
                                   bg_ratio=0.25), 3, 96),
    # head_dim 64 and a width that is a multiple of 128: the shapes at which the product takes its FUSED tcgen05
    # attention (csrc/flash_attn.cu), LayerNorm (csrc/layernorm.cu) and FFN paths -- the smaller cases above fall
    # to the composed library path.  The CVAE encoder runs under the `is_pad` key-padding mask in both.
    "maniskill_h128": (False, dict(hidden_dim=128, nhead=2, dim_feedforward=32, enc_layers=2, dec_layers=3,
                                   dropout=0.0, num_queries=12, action_dim=7, qpos_dim=9, goal_cond_dim=3,
                                   latent_dim=32, kl_weight=10.0, pcd_npoints=160, pcd_nsample=16), 3, 288),
    "rlbench_h128": (True, dict(hidden_dim=128, nhead=2, dim_feedforward=32, enc_layers=1, dec_layers=2,
                                dropout=0.0, num_queries=9, action_dim=11, qpos_dim=4, goal_cond_dim=16,
                                latent_dim=32, kl_weight=10.0, pcd_npoints=136, pcd_nsample=16, collision=True,
                                position_loss_weight=3.0), 2, 200),
}


def synth_batch(cfg, b, n, seed, ragged=True):
    """Batch contract of pcd_collate_fn (src/utils/sparse_tensor_utils.py:65-82), SURVEY.md 8b/8d."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(int(0.75 * n), n + 1, (b,), generator=g) if ragged else torch.full((b,), n)
    total = int(sizes.sum())
    coord = torch.rand(total, 3, generator=g) - 0.5
    color = torch.randint(0, 256, (total, 3), generator=g).float() / 127.5 - 1
    grid = torch.floor(coord / 0.005).long()
    grid = grid - grid.min(0).values
    a = cfg["action_dim"]
    actions = torch.randn(b, cfg["num_queries"], a, generator=g)
    if a > 7:  # rlbench: gripper / collision targets in [0, 1]
        actions[..., -2:] = torch.rand(b, cfg["num_queries"], 2, generator=g)
    npad = torch.randint(0, cfg["num_queries"] // 2 + 1, (b,), generator=g)
    is_pad = torch.arange(cfg["num_queries"])[None, :] >= (cfg["num_queries"] - npad)[:, None]
    batch = {
        "pcds": {"coord": coord, "grid_coord": grid, "feat": torch.cat([color, coord], 1),
                 "offset": torch.cumsum(sizes, 0)},
        "qpos": torch.randn(b, cfg["qpos_dim"], generator=g),
        "actions": actions, "is_pad": is_pad,
        "goal_cond": torch.randn(b, cfg["goal_cond_dim"], generator=g),
    }
    # foreground mask (use_mask data, rlbench_single_task_act.py:289-309): ~55% of every cloud, drawn
    # LAST so that the other tensors of the older fixtures are unchanged
    batch["pcds"]["mask"] = torch.rand(total, generator=g) < 0.55
    return batch


def clone_batch(batch):
    return {k: ({kk: vv.clone() for kk, vv in v.items()} if isinstance(v, dict) else v.clone()) for k, v in batch.items()}


def main():
    from .act_oracle import OraclePointNet

    act, tr, loss = install_reference_shim()
    OUT.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])  # optional: regenerate just the named cases
    for name, (rlbench, cfg, b, n) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(2024)
        backbone = OraclePointNet(6, int(cfg.get("backbone_classes", 0)))
        transformer = tr.Transformer(d_model=cfg["hidden_dim"], nhead=cfg["nhead"], num_encoder_layers=cfg["enc_layers"],
                                     num_decoder_layers=cfg["dec_layers"], dim_feedforward=cfg["dim_feedforward"],
                                     dropout=cfg["dropout"], normalize_before=False, return_intermediate_dec=True)
        encoder = tr.TransformerEncoder(d_model=cfg["hidden_dim"], nhead=cfg["nhead"], dim_feedforward=cfg["dim_feedforward"],
                                        dropout=cfg["dropout"], num_layers=cfg["enc_layers"], normalize_before=False)
        kw = dict(backbone=backbone, transformer=transformer, encoder=encoder, hidden_dim=cfg["hidden_dim"],
                  num_queries=cfg["num_queries"], num_cameras=1, action_dim=cfg["action_dim"], qpos_dim=cfg["qpos_dim"],
                  latent_dim=cfg["latent_dim"], action_loss=torch.nn.MSELoss(reduction="none"),
                  klloss=loss.KLDivergence(), kl_weight=cfg["kl_weight"], goal_cond_dim=cfg["goal_cond_dim"],
                  pcd_nsample=cfg["pcd_nsample"], pcd_npoints=cfg["pcd_npoints"])
        if cfg.get("pre_sample", 0) or cfg.get("use_mask", 0):
            kw.update(pre_sample=bool(cfg.get("pre_sample", 0)), use_mask=bool(cfg.get("use_mask", 0)),
                      bg_ratio=float(cfg.get("bg_ratio", 0.0)))
        if rlbench:
            model = act.ACTRLBenchPCD(**kw, collision=cfg["collision"], position_loss_weight=cfg["position_loss_weight"])
        else:
            model = act.ACTPCD(**kw)
        # de-trivialise parameters the default init leaves at 0 / 1
        with torch.no_grad():
            for pn, p in model.named_parameters():
                if pn.endswith("bias") or "norm" in pn or pn.endswith("bn.weight") or ".1.weight" in pn:
                    p.add_(0.1 * torch.randn_like(p))
            # parameters are rounded to fp16-representable values so the fixture stores them in
            # half the bytes without any loss (the model then RUNS in fp32 on exactly these values)
            for p in model.parameters():
                p.copy_(p.half().float())
        model.train()
        batch = synth_batch(cfg, b, n, seed=77)
        state = {k: v.detach().clone().numpy() for k, v in model.state_dict().items()}
        # inference branch (act.py:177-182,785-795): no actions -> latent 0, BatchNorm on the (initial) running
        # statistics, RLBench head converts rot6d -> quaternion.  Run BEFORE the training forward touches the buffers.
        model.eval()
        with torch.no_grad():
            ev = model({k: v for k, v in clone_batch(batch).items() if k in ("pcds", "qpos", "goal_cond")})
        eval_a_hat = ev["a_hat"].detach().numpy()
        model.train()
        torch.manual_seed(99)  # -> reparametrize eps
        out = model(clone_batch(batch))
        out["loss"].backward()
        torch.manual_seed(99)
        eps = torch.empty(b, cfg["latent_dim"]).normal_()
        # gradients are summarised per tensor (L2 norm, sum, 16 strided samples) to keep the fixture small
        grads = {}
        for k, p in model.named_parameters():
            if p.grad is not None:
                gflat = p.grad.detach().double().flatten()
                step = max(1, gflat.numel() // 16)
                grads[k] = np.concatenate([[gflat.norm().item(), gflat.sum().item()], gflat[::step][:16].numpy()])
        nograd = sorted(k for k, p in model.named_parameters() if p.grad is None)
        post = {k: v.detach().numpy() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
        flat = {}
        for k, v in state.items():
            flat["state/" + k] = v.astype(np.float16) if (v.dtype == np.float32 and "running" not in k and k != "pos_table") else v
        for k, v in grads.items():
            flat["grad/" + k] = v
        for k, v in post.items():
            flat["post/" + k] = v
        for k in ("coord", "grid_coord", "feat", "offset") + (("mask",) if cfg.get("use_mask", 0) else ()):
            flat["in/pcds/" + k] = batch["pcds"][k].numpy()
        for k in ("qpos", "actions", "is_pad", "goal_cond"):
            flat["in/" + k] = batch[k].numpy()
        flat["in/eps"] = eps.numpy()
        flat["eval/a_hat"] = eval_a_hat
        for k in ("a_hat", "is_pad_hat", "mu", "logvar", "loss", "action_loss", "kl_loss"):
            flat["out/" + k] = out[k].detach().numpy()
        flat["meta/nograd"] = np.array(nograd)
        flat["meta/cfg_keys"] = np.array(list(cfg.keys()))
        flat["meta/cfg_vals"] = np.array([float(v) for v in cfg.values()])
        np.savez_compressed(OUT / f"act_{name}.npz", **flat)
        print(name, "loss", float(out["loss"]), "params", sum(p.numel() for p in model.parameters()),
              "nograd", nograd, "file KB", (OUT / f"act_{name}.npz").stat().st_size // 1024)


if __name__ == "__main__":
    main()
