import sys, traceback
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from pointcloudmatters_b200.act import build_policy
from pointcloudmatters_b200.bc_module import ACTBCModule
from pointcloudmatters_b200.data import synthetic_act_batch, to_device
from pointcloudmatters_b200 import functional as PF

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = dict(bench.CFG2); 
policy = build_policy(cfg).to(dev).train()
module = ACTBCModule(policy, total_steps=1000)
module.configure_optimizers()
hb = synthetic_act_batch(8, 1024, seed=1)
b = to_device(hb, dev); b["pcds"]["n_max"] = hb["pcds"]["n_max"]
for i in range(3):
    module.training_step(b, i)
torch.cuda.synchronize()
tr = module._trainer

def attempt(name, fn, mode="global"):
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("OK  ", name, mode)
    except Exception as e:
        print("FAIL", name, mode, str(e).splitlines()[0][:150])
        torch.cuda.synchronize()

def clone_b():
    bb = tr._inputs_only(b)
    return {k: ({kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in v.items()} if isinstance(v, dict) else v.clone()) for k, v in bb.items()}

sb = clone_b()
with torch.no_grad():
    attempt("backbone fwd", lambda: policy.backbone(sb["pcds"]))
    attempt("fps+knn", lambda: policy.pcd_sampling((sb["pcds"]["coord"], torch.zeros(sb["pcds"]["coord"].shape[0], 512, device=dev), sb["pcds"]["offset"]), n_max=1024))
    attempt("full fwd nograd", lambda: policy(clone_b()) if False else policy(dict(sb, pcds=dict(sb["pcds"]))))
attempt("fwd+bwd global", lambda: tr._forward_backward(dict(sb, pcds=dict(sb["pcds"]))), "global")
attempt("fwd+bwd thread_local", lambda: tr._forward_backward(dict(sb, pcds=dict(sb["pcds"]))), "thread_local")
attempt("fwd+bwd relaxed", lambda: tr._forward_backward(dict(sb, pcds=dict(sb["pcds"]))), "relaxed")
