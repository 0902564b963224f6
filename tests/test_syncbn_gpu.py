"""SyncBatchNorm mode (configs/trainer/ddp.yaml:9) of the fused BatchNorm and set-abstraction operators: two ranks (two
processes sharing cuda:0, gloo backend for the tiny statistics collectives) each holding half of the rows must reproduce
the single-process result on the concatenated batch -- outputs, input gradients, running statistics -- and their summed
affine gradients must equal the single-process ones (torch.nn.SyncBatchNorm semantics)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _case(seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(700, 64, generator=g) * 2 + 0.5
    w = torch.randn(700, 64, generator=g)
    # set abstraction: 2 clouds x 160 points, 40 queries per cloud, 8 neighbours
    p = torch.rand(320, 3, generator=g)
    feat = torch.randn(320, 16, generator=g)
    wl = torch.randn(32, 19, generator=g) * 0.3
    q_idx = torch.cat([torch.randperm(160, generator=g)[:40], 160 + torch.randperm(160, generator=g)[:40]])
    nbr = torch.cat([torch.randint(0, 160, (40, 8), generator=g), 160 + torch.randint(0, 160, (40, 8), generator=g)]).int()
    wo = torch.randn(80, 32, generator=g)
    return x, w, p, feat, wl, q_idx, nbr, wo


def _run(rank, world, x, w, p, feat, wl, q_idx, nbr, wo):
    from pointcloudmatters_b200 import functional as PF

    dev = "cuda"
    rows = slice(rank * 350, (rank + 1) * 350) if world > 1 else slice(None)
    bn = torch.nn.BatchNorm1d(64).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(torch.linspace(0.5, 1.5, 64)); bn.bias.copy_(torch.linspace(-0.2, 0.2, 64))
    xi = x[rows].to(dev).requires_grad_(True)
    y = PF.batchnorm_relu(xi, bn)
    (y * w[rows].to(dev)).sum().backward()
    out = {"bn_y": y.detach().cpu(), "bn_dx": xi.grad.cpu(), "bn_dg": bn.weight.grad.cpu(), "bn_db": bn.bias.grad.cpu(),
           "bn_rm": bn.running_mean.cpu(), "bn_rv": bn.running_var.cpu()}
    # set abstraction on this rank's cloud(s)
    if world > 1:
        c0 = rank * 160
        ps, fs = p[c0:c0 + 160], feat[c0:c0 + 160]
        qi, nb, wos = q_idx[rank * 40:(rank + 1) * 40] - c0, nbr[rank * 40:(rank + 1) * 40] - c0, wo[rank * 40:(rank + 1) * 40]
    else:
        ps, fs, qi, nb, wos = p, feat, q_idx, nbr, wo
    lin = torch.nn.Linear(19, 32, bias=False).to(dev)
    sbn = torch.nn.BatchNorm1d(32).to(dev).train()
    with torch.no_grad():
        lin.weight.copy_(wl)
        sbn.weight.copy_(torch.linspace(0.6, 1.4, 32)); sbn.bias.copy_(torch.linspace(-0.1, 0.1, 32))
    fd = fs.to(dev).requires_grad_(True)
    pd = ps.to(dev)
    o = PF.set_abstraction(pd, fd, None, pd[qi.to(dev)].contiguous(), None, nb.to(dev), lin.weight, sbn)
    (o * wos.to(dev)).sum().backward()
    out.update(sa_y=o.detach().cpu(), sa_df=fd.grad.cpu(), sa_dw=lin.weight.grad.cpu(), sa_dg=sbn.weight.grad.cpu(),
               sa_db=sbn.bias.grad.cpu(), sa_rm=sbn.running_mean.cpu(), sa_rv=sbn.running_var.cpu())
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pointcloudmatters_b200 import functional as PF

    PF.set_sync_batchnorm(True)
    q.put((rank, _run(rank, world, *_case())))
    dist.destroy_process_group()


def test_syncbn_two_ranks_equal_one_process_on_the_concatenated_batch():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=300) for _ in range(2))
    [p.join(60) for p in procs]
    want = _run(0, 1, *_case())
    r0, r1 = res[0], res[1]
    close = lambda a, b, tol=2e-3: torch.testing.assert_close(a, b, rtol=tol, atol=tol * float(b.abs().max()) + 1e-6)
    for pre, n in (("bn", 350), ("sa", 40)):
        close(torch.cat([r0[pre + "_y"], r1[pre + "_y"]]), want[pre + "_y"], 2e-2 if pre == "sa" else 2e-4)
        close(r0[pre + "_dg"] + r1[pre + "_dg"], want[pre + "_dg"], 3e-2 if pre == "sa" else 2e-3)
        close(r0[pre + "_db"] + r1[pre + "_db"], want[pre + "_db"], 3e-2 if pre == "sa" else 2e-3)
        for rr in (r0, r1):  # identical running statistics on every rank = the single-process ones
            close(rr[pre + "_rm"], want[pre + "_rm"], 2e-2 if pre == "sa" else 2e-4)
            close(rr[pre + "_rv"], want[pre + "_rv"], 2e-2 if pre == "sa" else 2e-4)
    close(torch.cat([r0["bn_dx"], r1["bn_dx"]]), want["bn_dx"], 2e-3)
    close(torch.cat([r0["sa_df"], r1["sa_df"]]), want["sa_df"], 4e-2)
    close(r0["sa_dw"] + r1["sa_dw"], want["sa_dw"], 4e-2)
    # local statistics (the default) differ from the global ones on this data: the switch is doing something
    assert float((r0["bn_rm"] - _run_local_mean()).abs().max()) > 1e-3


def _run_local_mean():
    x = _case()[0]
    return 0.1 * x[:350].mean(0)
