"""Fused set-abstraction operator (csrc/sa_fused.cu) against a plain PyTorch fp32 reference of the
AS-WRITTEN chain of the reference (grouping -> Linear -> BatchNorm1d(train) -> ReLU -> max),
forward and backward, incl. -1-padded neighbourhoods and negative BatchNorm scales.
Tolerance: the Pf GEMM runs on bf16 operands (rel 4e-3 on pre-BN activations); everything else
is fp32 => outputs rel-L2 <= 1e-2, gradients rel-L2 <= 2.5e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pointops_oracle as O
from tests._data import clouds

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


def _reference(p, feat, new_p, idx, W, bn):
    m, k = idx.shape
    xyz_p = torch.cat([p, torch.zeros(1, 3, device=p.device)], 0)
    feat_p = torch.cat([feat, torch.zeros(1, feat.shape[1], device=p.device)], 0)
    flat = idx.reshape(-1).long()
    g = torch.cat([(xyz_p[flat].view(m, k, 3) - new_p.unsqueeze(1)) * torch.sign(idx + 1).unsqueeze(-1),
                   feat_p[flat].view(m, k, -1)], -1)
    y = F.relu(bn(F.linear(g, W).transpose(1, 2).contiguous()))
    return y.max(dim=-1).values


@pytest.mark.parametrize("b,n,mm,k,C,H,neg", [(3, 300, 64, 16, 64, 128, False), (2, 12, 6, 16, 32, 64, True),
                                               (2, 500, 100, 8, 512, 512, True)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("cloud_slices", [False, True])
def test_sa_fused_matches_as_written_chain(b, n, mm, k, C, H, neg, training, cloud_slices):
    """`cloud_slices`: the caller hands over the cloud offsets and a host-known size bound -> the kernels that keep a
    cloud's Pf slice in shared memory (k = 16 only; asserted through the per-entry-point call counters)."""
    from pointcloudmatters_b200 import functional as PF
    from pointcloudmatters_b200._lib import lib

    xyz, off, noff = clouds(b, n, mm, seed=3, ragged=True)
    fidx = O.farthest_point_sampling(xyz, off, noff)
    kidx, _ = O.knn_query(k, xyz, off, xyz[fidx], noff)
    g = torch.Generator().manual_seed(0)
    p = torch.from_numpy(xyz).cuda()
    new_p = p[torch.from_numpy(fidx).long().cuda()].contiguous()
    idx = torch.from_numpy(kidx).cuda()
    # bf16-representable inputs/weights so that the only difference is accumulation order
    feat0 = torch.randn(xyz.shape[0], C, generator=g).bfloat16().float().cuda()
    W0 = (torch.randn(H, 3 + C, generator=g) / np.sqrt(C)).cuda()
    W0[:, 3:] = W0[:, 3:].bfloat16().float()
    gamma = (1 + 0.3 * torch.randn(H, generator=g)).cuda()
    if neg:
        gamma[::3] *= -1  # negative scale -> the MIN over neighbours is selected
    beta = (0.2 * torch.randn(H, generator=g)).cuda()
    dout = torch.randn(new_p.shape[0], H, generator=g).cuda()

    outs = []
    for fused in (False, True):
        feat = feat0.clone().requires_grad_(True)
        W = W0.clone().requires_grad_(True)
        bn = torch.nn.BatchNorm1d(H).cuda()
        with torch.no_grad():
            bn.weight.copy_(gamma); bn.bias.copy_(beta)
            bn.running_mean.normal_(generator=None).mul_(0.1); bn.running_var.fill_(1.3)
        torch.manual_seed(1)
        bn.running_mean.copy_(torch.linspace(-0.2, 0.2, H)); 
        bn.train(training)
        if fused:
            before = dict(lib.calls)
            if cloud_slices:
                t_off, t_noff = torch.from_numpy(off).cuda(), torch.from_numpy(noff).cuda()
                out = PF.set_abstraction(p, feat, t_off, new_p, t_noff, idx, W, bn, n_max=int(np.diff(off, prepend=0).max()) + 5)
            else:
                out = PF.set_abstraction(p, feat, None, new_p, None, idx, W, bn)
            used = {n: lib.calls.get(n, 0) - before.get(n, 0) for n in ("pcm_sa_gather_stats", "pcm_sa_gather_sel", "pcm_sa_gather_sel_clouds")}
            if k == 16:  # single-extreme kernels (sign of the BatchNorm scale known up front), cloud-slice form when offsets are given
                assert used["pcm_sa_gather_stats"] == 0 and used["pcm_sa_gather_sel_clouds" if cloud_slices else "pcm_sa_gather_sel"] == 1, used
            else:
                assert used["pcm_sa_gather_stats"] == 1, used
        else:
            out = _reference(p, feat, new_p, idx, W, bn)
        out.backward(dout)
        outs.append((out.detach(), feat.grad, W.grad, bn.weight.grad, bn.bias.grad, bn.running_mean.clone(),
                     bn.running_var.clone(), int(bn.num_batches_tracked)))
    ref, got = outs
    assert _rel(got[0], ref[0]) <= 1e-2
    for i, name in [(1, "dfeat"), (2, "dW"), (3, "dgamma"), (4, "dbeta")]:
        assert _rel(got[i], ref[i]) <= 2.5e-2, (name, _rel(got[i], ref[i]))
    assert _rel(got[2][:, :3], ref[2][:, :3]) <= 2.5e-2  # the xyz columns separately
    torch.testing.assert_close(got[5], ref[5], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(got[6], ref[6], rtol=1e-3, atol=1e-4)
    assert got[7] == ref[7]
