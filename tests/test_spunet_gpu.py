"""SpUNet on the B200 path (csrc/spconv.cu rules + gathers, tcgen05 GEMMs, fused PDBatchNorm) against the dense-voxel
oracle (oracle/spunet_oracle.py).  PARITY UNPINNED: spconv, the library the reference builds on, is absent from the
reference tree and from this image; the oracle is pinned only to the stated sparse definitions
(tests/test_spunet_oracle_cpu.py).  Rule tables are integer work and compared exactly with brute-force dictionaries;
features use the bf16-operand tolerances of the other dense blocks."""
import itertools

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(seed, n=700, ext=14, batches=3):
    g = torch.Generator().manual_seed(seed)
    c = torch.unique(torch.cat([torch.randint(0, batches, (n, 1), generator=g), torch.randint(0, ext, (n, 3), generator=g)], 1), dim=0)
    return c, g


def test_rule_tables_are_exact():
    from pointcloudmatters_b200.spunet import SparseLevels

    coords, _ = _cloud(1)
    n = coords.shape[0]
    lv = SparseLevels(coords.int().cuda())
    lut = {tuple(c): i for i, c in enumerate(coords.tolist())}
    for k in (3, 5):
        nbr = lv.subm(0, k).cpu().numpy()
        want = np.full((n, k ** 3), -1, dtype=np.int32)
        for i, c in enumerate(coords.tolist()):
            for o, (a, b, cc) in enumerate(itertools.product(range(k), repeat=3)):
                want[i, o] = lut.get((c[0], c[1] + a - k // 2, c[2] + b - k // 2, c[3] + cc - k // 2), -1)
        assert np.array_equal(nbr, want)
    parent, kidx, child = (t.cpu().numpy() for t in lv.down(0))
    coarse = lv.coords[1].cpu().numpy()
    first = {}
    for i, c in enumerate(coords.tolist()):
        first.setdefault((c[0], c[1] // 2, c[2] // 2, c[3] // 2), i)
    order = sorted(first, key=first.get)  # coarse voxels numbered by their smallest child row
    assert coarse.shape[0] == len(order) and [tuple(r) for r in coarse.tolist()] == order
    idx = {k_: m for m, k_ in enumerate(order)}
    for i, c in enumerate(coords.tolist()):
        assert parent[i] == idx[(c[0], c[1] // 2, c[2] // 2, c[3] // 2)]
        assert kidx[i] == ((c[1] & 1) * 2 + (c[2] & 1)) * 2 + (c[3] & 1)
        assert child[parent[i], kidx[i]] == i
    assert (child >= 0).sum() == n
    # second level from the first
    p2, k2, c2 = (t.cpu().numpy() for t in lv.down(1))
    assert lv.coords[2].shape[0] == len({(c[0], c[1] // 4, c[2] // 4, c[3] // 4) for c in coords.tolist()})


def _pair(seed=0, **kw):
    from oracle.spunet_oracle import OracleSpUNet
    from pointcloudmatters_b200.spunet import SpUNet

    torch.manual_seed(seed)
    o = OracleSpUNet(**kw).train()
    with torch.no_grad():  # de-trivialise: BN affine / biases away from 1 / 0, weights larger than the 0.02 init
        for k, p in o.named_parameters():
            if "bns" in k or k.endswith("bias"):
                p.add_(0.2 * torch.randn_like(p))
            elif k.endswith("conv.weight") or "conv1" in k or "conv2" in k or "proj_conv" in k:
                p.mul_(8.0)
    m = SpUNet(**kw).cuda().train()
    m.load_state_dict(o.state_dict())
    return o, m


@pytest.mark.parametrize("cls_mode", [False, True])
def test_spunet_matches_dense_oracle(cls_mode):
    kw = dict(in_channels=6, num_classes=24 if not cls_mode else 0, base_channels=16, channels=(16, 32, 48, 64, 64, 48, 32, 32),
              layers=(1, 2, 1, 1, 1, 1, 1, 1), cls_mode=cls_mode)
    if cls_mode:  # the classification head averages the DEEPEST level: two stages keep a few hundred rows there (four leave 22,
        kw.update(channels=(16, 32, 32, 32), layers=(1, 2, 1, 1))  # where training-mode BatchNorm makes gradients ill-conditioned)
    o, m = _pair(3, **kw)
    coords, g = _cloud(2, n=900, ext=18, batches=3)
    n = coords.shape[0]
    sizes = torch.bincount(coords[:, 0], minlength=3)
    feat = torch.randn(n, 6, generator=g)
    inp = dict(grid_coord=coords[:, 1:].contiguous(), feat=feat, offset=torch.cumsum(sizes, 0), condition=["S3DIS"])
    want = o(dict(inp))
    (0.5 * want.pow(2).sum()).backward()
    got = m({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items()})
    assert got.shape == want.shape
    rel = float((got.detach().cpu() - want.detach()).norm() / want.detach().norm())
    assert rel <= 3e-2, rel
    (0.5 * got.pow(2).sum()).backward()
    # Gradients of a quadratic loss, per tensor, relative to max(own norm, 5 % of the largest gradient norm): the FiLM /
    # affine parameters of a BatchNorm that feeds conv -> BatchNorm have a (near-)zero true gradient (scale invariance),
    # so their own norm is no yardstick.  Measured on the CPU: rounding the ORACLE's conv operands to bf16 moves its
    # gradients by <= 8e-2 (worst: *.modulation.1.weight) with a median of 1e-3 under this metric; every layer type alone is
    # held to 1.5e-2 and PDBatchNorm to 1e-3 in the tests below.
    og = dict(o.named_parameters())
    scale = max(float(p.grad.norm()) for p in og.values() if p.grad is not None)
    errs = {}
    for k, p in m.named_parameters():
        if og[k].grad is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        errs[k] = float((p.grad.cpu() - og[k].grad).norm()) / max(float(og[k].grad.norm()), 5e-2 * scale)
    # *.modulation.1.weight (FiLM of a BatchNorm that feeds conv -> BatchNorm): (near-)zero true gradient, see above; the op
    # itself is pinned at 1e-3 in test_pdbatchnorm_matches_reference_code_path
    bad = {k: v for k, v in errs.items() if v > (3.5e-1 if k.endswith("modulation.1.weight") else 1.5e-1)}
    assert not bad, bad
    assert sorted(errs.values())[len(errs) // 2] <= 2e-2
    # running statistics of ALL per-condition BatchNorm copies were updated, exactly like the reference's loop
    so, sm = o.state_dict(), m.state_dict()
    for k in so:
        if "running" in k:
            torch.testing.assert_close(sm[k].cpu(), so[k], rtol=2e-2, atol=2e-3, msg=k)
        if "num_batches_tracked" in k:
            assert int(sm[k]) == int(so[k]) == 1, k


def test_spunet_backbone_drives_the_diffusion_policy_encoder():
    """BASELINE cfg-3's encoder as written: SpUNet(num_classes=96) behind PCDObsEncoder; one compute_loss + backward."""
    from pointcloudmatters_b200.data import synthetic_dp_batch, to_device
    from pointcloudmatters_b200.diffusion import DDPMScheduler, DiffusionUnetImagePolicy, PCDObsEncoder
    from pointcloudmatters_b200.spunet import SpUNet

    torch.manual_seed(0)
    shape_meta = {"obs": {"pcds": {"shape": [6], "type": "pcd"}, "qpos": {"shape": [9], "type": "low_dim"}}, "action": {"shape": [7]},
                  "goal": None}
    backbone = SpUNet(6, num_classes=32, base_channels=16, channels=(16, 32, 32, 32, 32, 32, 32, 32), layers=(1, 1, 1, 1, 1, 1, 1, 1))
    enc = PCDObsEncoder(shape_meta, backbone, share_pcd_model=True, n_obs_step=2, pcd_nsample=16, pcd_npoints=64, pcd_hidden_dim=32,
                        projector_layers=1, projector_channels=[32, 64, 64])
    policy = DiffusionUnetImagePolicy(shape_meta, DDPMScheduler(num_train_timesteps=100), enc, horizon=16, n_action_steps=8, n_obs_steps=2,
                                      diffusion_step_embed_dim=64, down_dims=[64, 128], kernel_size=5, n_groups=8).cuda().train()
    policy.normalizer.set_identity({"qpos": 9, "action": 7}).to("cuda")
    hb = synthetic_dp_batch(4, 300, seed=2, ragged=True)
    # voxelise on a coarse grid so that neighbouring voxels exist (0.005 m cells of uniform noise would all be isolated)
    hb["obs"]["pcds"]["grid_coord"] = torch.floor((hb["obs"]["pcds"]["coord"] + 0.5) / 0.06).long()
    b = to_device(hb, "cuda")
    b["obs"]["pcds"]["n_max"] = hb["obs"]["pcds"]["n_max"]
    loss = policy.compute_loss(b)["loss"]
    loss.backward()
    assert float(loss) == float(loss)
    g = backbone.conv_input.conv.weight.grad
    assert g is not None and float(g.abs().sum()) > 0


@pytest.mark.parametrize("kind,cin,cout,k", [("subm", 16, 24, 3), ("subm", 6, 16, 5), ("down", 16, 32, 2), ("inverse", 32, 16, 2)])
def test_sparse_convolution_ops_match_dense_oracle(kind, cin, cout, k):
    """Each layer type alone, forward and all three gradients, against the oracle's dense convolution."""
    from oracle import spunet_oracle as O
    from pointcloudmatters_b200 import spunet as P

    coords, g = _cloud(5, n=800, ext=16, batches=2)
    lv = P.SparseLevels(coords.int().cuda())
    x0 = O.Sp(torch.randn(coords.shape[0], cin if kind != "inverse" else 8, generator=g), coords.long())
    if kind == "inverse":  # needs a coarse input: go down first with the oracle, feed both sides the same coarse features
        down_o = O.SparseConv3d(8, cin, 2)
        xo = down_o(x0)
        xo = O.Sp(xo.features.detach().requires_grad_(True), xo.coords, xo.skip)
        lv.down(0)
        # the product numbers coarse voxels by smallest child row, the oracle by sorted coordinate: permute
        lut = {tuple(c): i for i, c in enumerate(xo.coords.tolist())}
        perm = torch.tensor([lut[tuple(c)] for c in lv.coords[1].cpu().tolist()])
        xp = P.SparseTensor(xo.features.detach()[perm].cuda().requires_grad_(True), lv, 1)
        co, cp = O.SparseInverseConv3d(cin, cout, 2), P.SparseInverseConv3d(cin, cout, 2).cuda()
    else:
        xo = O.Sp(x0.features.clone().requires_grad_(True), x0.coords)
        xp = P.SparseTensor(x0.features.cuda().requires_grad_(True), lv, 0)
        perm = None
        if kind == "subm":
            co, cp = O.SubMConv3d(cin, cout, k), P.SubMConv3d(cin, cout, k).cuda()
        else:
            co, cp = O.SparseConv3d(cin, cout, 2), P.SparseConv3d(cin, cout, 2).cuda()
    with torch.no_grad():
        co.weight.mul_(10.0)
        cp.weight.copy_(co.weight)
    yo, yp = co(xo), cp(xp)
    fo, fp = yo.features, yp.features
    if kind == "down":  # output rows: oracle sorted by coordinate, product by smallest child row
        lut = {tuple(c): i for i, c in enumerate(yo.coords.tolist())}
        perm_out = torch.tensor([lut[tuple(c)] for c in lv.coords[1].cpu().tolist()])
        fo = fo[perm_out]
    assert fo.shape == fp.shape
    assert float((fp.detach().cpu() - fo.detach()).norm() / fo.detach().norm()) <= 1e-2
    w = torch.randn(fo.shape, generator=g)
    (fo * w).sum().backward()
    (fp * w.cuda()).sum().backward()
    gxo = xo.features.grad if perm is None else xo.features.grad[perm]
    for name, a, b in (("dx", xp.features.grad.cpu(), gxo), ("dW", cp.weight.grad.cpu(), co.weight.grad)):
        err = float((a - b).norm() / b.norm())
        assert err <= 1.5e-2, (name, err)


def test_pdbatchnorm_matches_reference_code_path():
    from oracle import spunet_oracle as O
    from pointcloudmatters_b200 import spunet as P

    torch.manual_seed(1)
    for adaptive in (False, True):
        bo = O.PDBatchNorm(32, adaptive=adaptive).train()
        with torch.no_grad():
            for p in bo.parameters():
                p.add_(0.3 * torch.randn_like(p))
        bp = P.PDBatchNorm(32, adaptive=adaptive).cuda().train()
        bp.load_state_dict(bo.state_dict())
        x = torch.randn(500, 32) * 2 + 1
        ctx = torch.randn(1, 256)
        xo, xp = x.clone().requires_grad_(True), x.cuda().requires_grad_(True)
        co, cp = ctx.clone().requires_grad_(True), ctx.cuda().requires_grad_(True)
        yo = torch.relu(bo(xo, "S3DIS", co if adaptive else None))
        yp = bp(xp, "S3DIS", cp if adaptive else None, relu=True)
        torch.testing.assert_close(yp.detach().cpu(), yo.detach(), rtol=1e-4, atol=1e-4)
        w = torch.randn(500, 32)
        (yo * w).sum().backward()
        (yp * w.cuda()).sum().backward()
        torch.testing.assert_close(xp.grad.cpu(), xo.grad, rtol=1e-3, atol=1e-4)
        if adaptive:
            torch.testing.assert_close(cp.grad.cpu(), co.grad, rtol=1e-3, atol=1e-4)
        po = dict(bo.named_parameters())
        for k, p in bp.named_parameters():
            torch.testing.assert_close(p.grad.cpu(), po[k].grad, rtol=1e-3, atol=1e-4, msg=k)
        so = bo.state_dict()
        for k, v in bp.state_dict().items():
            torch.testing.assert_close(v.cpu(), so[k], rtol=1e-4, atol=1e-5, msg=k)
