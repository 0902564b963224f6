"""oracle/spunet_oracle.py (dense-voxel restatement of the reference SpUNet; PARITY UNPINNED -- spconv is absent) against
brute-force loops over the stated definitions of the three sparse layer types, and its state_dict surface against the
key list of the reference class (module tree of spunet.py:229-372 written out by hand below)."""
import itertools

import torch

from oracle.spunet_oracle import OracleSpUNet, Sp, SparseConv3d, SparseInverseConv3d, SubMConv3d


def _cloud(seed, n=260, ext=11, batches=2):
    g = torch.Generator().manual_seed(seed)
    c = torch.unique(torch.cat([torch.randint(0, batches, (n, 1), generator=g), torch.randint(0, ext, (n, 3), generator=g)], 1), dim=0)
    return c, g


def test_dense_convolutions_equal_their_sparse_definitions():
    coords, g = _cloud(0)
    n = coords.shape[0]
    x = Sp(torch.randn(n, 4, generator=g), coords)
    lut = {tuple(c.tolist()): i for i, c in enumerate(coords)}
    for k in (3, 5):
        conv = SubMConv3d(4, 5, k)
        got = conv(x).features
        want = torch.zeros(n, 5)
        for i, c in enumerate(coords.tolist()):
            for a, b, cc in itertools.product(range(k), repeat=3):
                j = lut.get((c[0], c[1] + a - k // 2, c[2] + b - k // 2, c[3] + cc - k // 2))
                if j is not None:
                    want[i] += conv.weight[:, a, b, cc, :] @ x.features[j]
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    d = SparseConv3d(4, 5, 2)
    y = d(x)
    lutc = {tuple(c.tolist()): i for i, c in enumerate(y.coords)}
    assert len(lutc) == len({(c[0], c[1] // 2, c[2] // 2, c[3] // 2) for c in coords.tolist()})
    want = torch.zeros(y.coords.shape[0], 5)
    for i, c in enumerate(coords.tolist()):
        want[lutc[(c[0], c[1] // 2, c[2] // 2, c[3] // 2)]] += d.weight[:, c[1] % 2, c[2] % 2, c[3] % 2, :] @ x.features[i]
    torch.testing.assert_close(y.features, want, rtol=1e-5, atol=1e-6)
    u = SparseInverseConv3d(5, 3, 2)
    z = u(y)
    want = torch.zeros(n, 3)
    for i, c in enumerate(coords.tolist()):
        want[i] = u.weight[:, c[1] % 2, c[2] % 2, c[3] % 2, :] @ y.features[lutc[(c[0], c[1] // 2, c[2] // 2, c[3] // 2)]]
    torch.testing.assert_close(z.features, want, rtol=1e-5, atol=1e-6)
    assert torch.equal(z.coords, coords)


def test_state_dict_surface_follows_the_reference_module_tree():
    m = OracleSpUNet(6, num_classes=96)  # reference defaults: channels (32..96), layers (2,3,4,6,2,2,2,2)
    keys = set(m.state_dict())
    bn = lambda p: {f"{p}.bns.{i}.{s}" for i in range(3) for s in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")} | \
        {f"{p}.modulation.1.weight", f"{p}.modulation.1.bias"}
    want = {"embedding_table.weight", "conv_input.conv.weight", "final.weight", "final.bias"} | bn("conv_input.bn")
    ch, ly = (32, 64, 128, 256, 256, 128, 96, 96), (2, 3, 4, 6, 2, 2, 2, 2)
    enc_c, dec_c = 32, ch[-1]
    for s in range(4):
        want |= {f"down.{s}.conv.weight", f"up.{s}.conv.weight"} | bn(f"down.{s}.bn") | bn(f"up.{s}.bn")
        for i in range(ly[s]):
            p = f"enc.{s}.block{i}"
            want |= {f"{p}.conv1.weight", f"{p}.conv2.weight"} | bn(f"{p}.bn1") | bn(f"{p}.bn2")
        for i in range(ly[len(ch) - s - 1]):
            p = f"dec.{s}.block{i}"
            want |= {f"{p}.conv1.weight", f"{p}.conv2.weight"} | bn(f"{p}.bn1") | bn(f"{p}.bn2")
            if i == 0:  # in_channels = dec + enc != embed_channels -> projection (spunet.py:97-103)
                want |= {f"{p}.proj_conv.weight"} | bn(f"{p}.proj_norm")
        enc_c, dec_c = ch[s], ch[len(ch) - s - 2]
    assert keys == want
    sd = m.state_dict()
    assert tuple(sd["conv_input.conv.weight"].shape) == (32, 5, 5, 5, 6)      # spconv 2.x layout (out, kD, kH, kW, in)
    assert tuple(sd["down.1.conv.weight"].shape) == (64, 2, 2, 2, 32)
    assert tuple(sd["up.3.conv.weight"].shape) == (256, 2, 2, 2, 256)
    assert tuple(sd["dec.0.block0.proj_conv.weight"].shape) == (96, 1, 1, 1, 128)
    assert tuple(sd["final.weight"].shape) == (96, 1, 1, 1, 96)


def test_product_module_has_the_same_state_dict_surface():
    from pointcloudmatters_b200.spunet import SpUNet

    kw = dict(in_channels=6, num_classes=96)
    a, b = OracleSpUNet(**kw).state_dict(), SpUNet(**kw).state_dict()
    assert {k: tuple(v.shape) for k, v in a.items()} == {k: tuple(v.shape) for k, v in b.items()}
