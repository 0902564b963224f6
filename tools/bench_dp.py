#!/usr/bin/env python
"""tools/bench_dp.py -- Diffusion-Policy training step (BASELINE cfg-3 shapes
with the PointNet backbone: N = 1024 points, M = 512, 2 obs steps, horizon 16, U-Net [512, 1024, 2048], 255.8 M
parameters) on ONE B200.  Not the headline bench (bench.py = cfg-2); this records the step time, the per-family
kernel times and the GEMM roofline of the second policy family for DESIGN.md / profiles/.

    python tools/bench_dp.py [--batch 16] [--steps 20] [--warmup 5] [--no-cuda-graph]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def infer(args):
    """Closed-loop inference latency: encoder + 100 denoising steps (`num_inference_steps: 100`,
    maniskill2_diffusion_policy_model.yaml:46), CUDA-graphed step vs eager launches."""
    import torch

    from pointcloudmatters_b200.data import synthetic_dp_batch, to_device
    from pointcloudmatters_b200.diffusion import DP_MODEL_CFG, build_dp_policy

    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    cfg = dict(DP_MODEL_CFG, pcd_npoints=args.points // 2, num_inference_steps=100)
    policy = build_dp_policy(cfg).to(dev).eval()
    policy.normalizer.set_identity({"qpos": cfg["qpos_dim"], "action": cfg["action_dim"]}).to(dev)
    h = synthetic_dp_batch(args.batch, args.points, seed=1)
    obs = to_device({"obs": h["obs"]}, dev)
    obs["obs"]["pcds"]["n_max"] = h["obs"]["pcds"]["n_max"]
    out = {}
    for graph in (True, False):
        for _ in range(2):
            policy.predict_action(obs, use_cuda_graph=graph)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            a = policy.predict_action(obs, use_cuda_graph=graph)
        e1.record()
        torch.cuda.synchronize()
        out["graph" if graph else "eager"] = e0.elapsed_time(e1) / n
    print(json.dumps({"workload": f"DP predict_action, {args.batch} env(s), N={args.points}, 100 DDPM steps, 255.8 M-param U-Net",
                      "ms_per_call_cuda_graph": out["graph"], "ms_per_call_eager": out["eager"],
                      "ms_per_denoising_step_cuda_graph": out["graph"] / 100, "action_shape": list(a["action"].shape)}))


def main():
    import torch

    from pointcloudmatters_b200 import functional as PF
    from pointcloudmatters_b200.bc_module import DiffusionPolicyBCModule
    from pointcloudmatters_b200.data import synthetic_dp_batch, to_device
    from pointcloudmatters_b200.diffusion import DP_MODEL_CFG, build_dp_policy

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU (cfg-3: 128 global / 8 GPUs)")
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cuda-graph", action="store_true")
    ap.add_argument("--infer", action="store_true", help="time predict_action (100-step DDPM sampling) instead of training")
    args = ap.parse_args()
    if args.infer:
        return infer(args)
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    cfg = dict(DP_MODEL_CFG, pcd_npoints=args.points // 2)
    policy = build_dp_policy(cfg).to(dev).train()
    policy.normalizer.set_identity({"qpos": cfg["qpos_dim"], "action": cfg["action_dim"]}).to(dev)
    module = DiffusionPolicyBCModule(policy, total_steps=1000, use_cuda_graph=not args.no_cuda_graph)
    module.configure_optimizers()
    host = [synthetic_dp_batch(args.batch, args.points, seed=1000 + i) for i in range(4)]
    res = []
    for h in host:
        r = to_device(h, dev)
        r["obs"]["pcds"]["n_max"] = h["obs"]["pcds"]["n_max"]
        res.append(r)

    def run(steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            loss = module.training_step(res[i % len(res)], i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, float(loss)

    run(max(3, args.warmup) + 2)
    ms, loss = run(args.steps)
    tr = module._trainer
    graph_was, tr.use_cuda_graph = tr.use_cuda_graph, False
    run(1)
    PF.KERNEL_TIMER.start()
    for i in range(3):
        torch.cuda._sleep(int(0.06 * 1.9e9))
        module.training_step(res[i % len(res)], i)
    ks = PF.KERNEL_TIMER.summary()
    PF.KERNEL_TIMER.stop()
    PF.retime_gemm_shapes(ks)
    try:
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "dp_gemm_by_shape.json").write_text(json.dumps(ks, indent=1))
    except Exception:
        pass
    for v in ks.values():
        v.pop("by_shape", None)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    n_params = sum(p.numel() for p in policy.parameters())
    print(json.dumps({"workload": f"cfg3-shape DP step: PointNet(96)+SA+U-Net[512,1024,2048], N={args.points}, "
                                  f"{args.batch} samples ({2 * args.batch} clouds) on 1 GPU",
                      "ms_per_step": ms, "steps_per_sec": 1e3 / ms, "samples_per_sec": 1e3 / ms * args.batch,
                      "cuda_graph": bool(graph_was), "params_M": n_params / 1e6, "last_loss": loss,
                      "roofline": PF.roofline_for(ks, peaks, 3),
                      "kernel_ms_per_step": {k: v.get("total_ms_isolated", v["total_ms"]) / 3 for k, v in ks.items()},
                      "optimizer_hbm_floor_ms": n_params * 34 / (peaks.get("hbm_gbs", 6550.0) * 1e9) * 1e3}))


if __name__ == "__main__":
    main()
