#!/usr/bin/env python
"""R-GPU baseline: the REFERENCE's own GPU path on one B200.

    bash baseline/install_ref.sh                 # build container: installs the unmodified reference into baseline/_ref
    python tools/bench_ref_gpu.py [--steps 20 --warmup 5 --batch 64 --tf32 0|1]

What runs: the unmodified `pointops` package of the reference (its Python wrappers + its own pybind11 CUDA extension
`pointops._C`, compiled by its own setup.py for sm_100) and the reference's own `ACTPCD` / `Transformer` /
`TransformerEncoder` modules imported from baseline/_ref/src, fp32, dropout 0.1, one training step =
forward + backward + clip_grad_norm_(0.5) + torch.optim.AdamW(lr 5e-5, wd 0.05) -- on the cfg-2 batch of bench.py.
Nothing of pointcloudmatters_b200 is on this path (only its synthetic batch generator is shared).

Substitutions (the box has no lightning / hydra / spconv):
  * `src`, `src.utils`, `src.models...` package __init__ files import lightning / hydra: the packages are registered as
    empty namespace stubs so that only the submodules the policy needs are executed (act.py, transformer.py, utils.py,
    loss/misc.py, sparse_tensor_utils.py, rotation_conversions.py) -- all unmodified;
  * the spconv `PointNet` backbone (pointnet.py:16-85): k=1 SubMConv3d == row-wise Linear on unique voxels, restated
    below as Linear + BatchNorm1d(eps 1e-3, momentum 0.01) + ReLU (same FLOPs, no hash / indice-pair build, so this
    FAVOURS the reference).
Prints one JSON line and writes gpurun_out/ref_gpu_baseline.json.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import json
import os
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
sys.path.insert(0, str(ROOT))


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [str(path)]
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference():
    import torch

    if not (REF / "pointops").exists():
        raise SystemExit("baseline/_ref is empty: run `bash baseline/install_ref.sh` in the build container first")
    sys.path.insert(0, str(REF))
    try:  # legacy typed constructors the reference wrappers allocate with (functions/sampling.py:17-18)
        torch.cuda.IntTensor(1)
    except Exception:
        torch.cuda.IntTensor = lambda *s: torch.empty(*s, dtype=torch.int32, device="cuda")
        torch.cuda.FloatTensor = lambda *s: torch.empty(*s, dtype=torch.float32, device="cuda")
    import pointops  # noqa: F401  (the reference's own package + CUDA extension)

    src = REF / "src"
    _pkg("src", src)
    utils = _pkg("src.utils", src / "utils")
    stu = _load("src.utils.sparse_tensor_utils", src / "utils" / "sparse_tensor_utils.py")
    utils.offset2batch = stu.offset2batch
    _load("src.utils.rotation_conversions", src / "utils" / "rotation_conversions.py")
    for p in ("src.models", "src.models.components", "src.models.components.act", "src.models.components.loss"):
        _pkg(p, src.joinpath(*p.split(".")[1:]))
    act = importlib.import_module("src.models.components.act.act")
    tr = importlib.import_module("src.models.components.act.transformer")
    loss = importlib.import_module("src.models.components.loss.misc")
    return act, tr, loss, pointops


def make_pointnet(in_channels=6):
    import torch.nn as nn

    class LinearPointNet(nn.Module):
        """pointnet.py:16-85 with every SubMConv3d(k=1, bias=False) as nn.Linear(bias=False)."""

        def __init__(self):
            super().__init__()
            self.in_channels, self.num_classes, self.num_channels = in_channels, 0, 512
            dims = [in_channels, 64, 64, 64, 128, 512]
            self.layers = nn.ModuleList(
                nn.Sequential(nn.Linear(a, b, bias=False), nn.BatchNorm1d(b, eps=1e-3, momentum=0.01), nn.ReLU())
                for a, b in zip(dims[:-1], dims[1:]))

        def forward(self, input_dict):
            x = input_dict["feat"]
            for layer in self.layers:
                x = layer(x)
            return x

    return LinearPointNet()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tf32", type=int, default=0, help="1: allow TF32 matmuls (not the reference default)")
    ap.add_argument("--amp", type=int, default=0, help="1: torch.autocast(bf16) around forward (not the reference default)")
    ap.add_argument("--stage-times", type=int, default=1)
    args = ap.parse_args()
    import torch

    from bench import CFG2, N_POINTS, WORKLOAD
    from pointcloudmatters_b200.data import synthetic_act_batch, to_device

    act, tr, loss, pointops = import_reference()
    torch.backends.cuda.matmul.allow_tf32 = bool(args.tf32)
    torch.backends.cudnn.allow_tf32 = bool(args.tf32)
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    c = CFG2
    transformer = tr.Transformer(d_model=c["hidden_dim"], nhead=c["nhead"], num_encoder_layers=c["enc_layers"],
                                 num_decoder_layers=c["dec_layers"], dim_feedforward=c["dim_feedforward"], dropout=c["dropout"],
                                 normalize_before=False, return_intermediate_dec=True)
    encoder = tr.TransformerEncoder(d_model=c["hidden_dim"], nhead=c["nhead"], dim_feedforward=c["dim_feedforward"],
                                    dropout=c["dropout"], num_layers=c["enc_layers"], normalize_before=False)
    model = act.ACTPCD(backbone=make_pointnet(), transformer=transformer, encoder=encoder, hidden_dim=c["hidden_dim"],
                       num_queries=c["num_queries"], num_cameras=1, action_dim=c["action_dim"], qpos_dim=c["qpos_dim"],
                       latent_dim=c["latent_dim"], action_loss=torch.nn.MSELoss(reduction="none"), klloss=loss.KLDivergence(),
                       kl_weight=c["kl_weight"], goal_cond_dim=c["goal_cond_dim"], pcd_nsample=c["pcd_nsample"],
                       pcd_npoints=c["pcd_npoints"]).to(dev).train()
    opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.05)
    host = [synthetic_act_batch(args.batch, N_POINTS, seed=1000 + i) for i in range(4)]
    for h in host:
        h["pcds"].pop("n_max")
    batches = [to_device(h, dev) for h in host]

    def clone(b):
        return {k: (dict(v) if isinstance(v, dict) else v) for k, v in b.items()}

    stage = {}

    def timed(name, fn):
        if not args.stage_times:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        stage.setdefault(name, []).append((e0, e1))
        return out

    def step(i):
        b = clone(batches[i % len(batches)])
        opt.zero_grad(set_to_none=True)
        if args.amp:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = timed("forward", lambda: model(b))
        else:
            out = timed("forward", lambda: model(b))
        timed("backward", lambda: out["loss"].backward())
        timed("clip+adamw", lambda: (torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5), opt.step()))
        return out["loss"]

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    stage.clear()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    marks[0].record()
    for i in range(args.steps):
        last = step(i)
        marks[i + 1].record()
    torch.cuda.synchronize()
    per = sorted(a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:]))
    total = marks[0].elapsed_time(marks[-1])
    ms = total / args.steps
    # FPS / kNN kernels of the reference alone, same inputs (the ops the drop-in replaces)
    p, o = batches[0]["pcds"]["coord"], batches[0]["pcds"]["offset"]
    n_o = torch.arange(1, args.batch + 1, device=dev) * c["pcd_npoints"]

    def ev(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    fps_ms, idx = ev(lambda: pointops.farthest_point_sampling(p, o.int(), n_o.int()))
    n_p = p[idx.long()].contiguous()
    knn_ms, _ = ev(lambda: pointops.knn_query(c["pcd_nsample"], p, o.int(), n_p, n_o.int()))
    line = {"metric": "bc_train_steps_per_sec", "impl": "reference-gpu (R-GPU)", "value": 1e3 / ms, "unit": "steps/s",
            "ms_per_step": ms, "step_ms": {"median": per[len(per) // 2], "p10": per[len(per) // 10], "p90": per[(9 * len(per)) // 10]},
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "dtype": "bf16-autocast" if args.amp else ("tf32" if args.tf32 else "f32"),
            "config": {"workload": WORKLOAD, "global_batch": args.batch,
                       "path": "reference pointops (own CUDA extension, sm_100) + reference ACTPCD/Transformer modules, torch eager; "
                               "PointNet backbone = Linear restatement (spconv absent)"},
            "stage_ms": {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in stage.items()},
            "pointops_wrapper_ms": {"farthest_point_sampling (incl. its O(B) host syncs)": fps_ms, "knn_query": knn_ms},
            "last_loss": float(last.detach()), "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    tag = "amp" if args.amp else ("tf32" if args.tf32 else "fp32")
    (ROOT / "gpurun_out" / f"ref_gpu_baseline_{tag}.json").write_text(json.dumps(line, indent=1))
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    os.environ.setdefault("CUDA_MODULE_LOADING", "LAZY")
    main()
