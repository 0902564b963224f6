"""oracle/spunet_oracle.py -- TEST INFRASTRUCTURE (never imported by the product).

Dense-voxel CPU restatement of the reference SpUNet (src/models/components/pcd_encoder/spunet.py:19-463) for small grids.
PARITY UNPINNED: the reference computes its convolutions with the third-party `spconv` library (README.md:119-123
`pip3 install spconv-cu118`, not in requirements.txt, not vendored, absent from this image), so this oracle restates the
published semantics of its three layer types on a DENSIFIED grid with torch's own dense convolutions -- an
implementation that shares no code and no data structure with the product's rule tables:
  * SubMConv3d(k)            : F.conv3d(dense, W, padding=k//2), read back at the ACTIVE sites only
                               (submanifold: output set = input set; spunet.py:95-122,210-217,368-372);
  * SparseConv3d(k=2, s=2)   : F.conv3d(dense, W, stride=2); active outputs = { c // 2 } (spunet.py:159-166);
  * SparseInverseConv3d(k=2) : F.conv_transpose3d(dense_coarse, W, stride=2), read back at the fine active sites
                               (spunet.py:189-195);
weights in spconv 2.x layout (out, kD, kH, kW, in) <-> torch (out, in, kD, kH, kW).  `PDBatchNorm` is the reference's
own code path (every per-condition BatchNorm sees the input, the selected one is kept; FiLM modulation, :54-73).
Module / parameter names equal the reference's, so one state_dict drives oracle and product.
"""
from __future__ import annotations

from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F


class PDBatchNorm(nn.Module):
    def __init__(self, num_features, context_channels=256, eps=1e-3, momentum=0.01,
                 conditions=("ScanNet", "S3DIS", "Structured3D"), decouple=True, adaptive=False, affine=True):
        super().__init__()
        self.conditions, self.decouple, self.adaptive, self.affine = conditions, decouple, adaptive, affine
        if decouple:
            self.bns = nn.ModuleList([nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine) for _ in conditions])
        else:
            self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine)
        if adaptive:
            self.modulation = nn.Sequential(nn.SiLU(), nn.Linear(context_channels, 2 * num_features, bias=True))

    def forward(self, feat, condition=None, context=None):
        if self.decouple:
            assert condition in self.conditions
            _feat = 0
            for c, bn in zip(self.conditions, self.bns):
                _feat = _feat + bn(feat) * (1 if c == condition else 0)
            feat = _feat
        else:
            feat = self.bn(feat)
        if self.adaptive:
            shift, scale = self.modulation(context).chunk(2, dim=1)
            feat = feat * (1.0 + scale) + shift
        return feat


class Sp:
    """features (n, C) on active sites coords (n, 4) = [b, x, y, z]."""

    def __init__(self, features, coords, skip=None):
        self.features, self.coords, self.skip = features, coords, skip

    def replace_feature(self, f):
        return Sp(f, self.coords, self.skip)

    def dense(self, pad_even=False):
        b = int(self.coords[:, 0].max()) + 1
        ext = (self.coords[:, 1:].max(0).values + 1).tolist()
        if pad_even:
            ext = [e + (e & 1) for e in ext]
        d = torch.zeros((b, *ext, self.features.shape[1]), dtype=self.features.dtype)
        c = self.coords.long()
        return d.index_put((c[:, 0], c[:, 1], c[:, 2], c[:, 3]), self.features).permute(0, 4, 1, 2, 3)

    def read(self, dense):
        c = self.coords.long()
        return dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]


class _Conv(nn.Module):
    def __init__(self, cin, cout, k, bias=False, **_):
        super().__init__()
        self.k = k
        self.weight = nn.Parameter(torch.empty(cout, k, k, k, cin))
        nn.init.trunc_normal_(self.weight, std=0.02)
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None

    @property
    def w(self):
        return self.weight.permute(0, 4, 1, 2, 3)


class SubMConv3d(_Conv):
    def forward(self, x):
        return x.replace_feature(x.read(F.conv3d(x.dense(), self.w, self.bias, padding=self.k // 2)))


class SparseConv3d(_Conv):
    def forward(self, x):
        out = F.conv3d(x.dense(pad_even=True), self.w, self.bias, stride=2)
        coarse = torch.unique(torch.cat([x.coords[:, :1], x.coords[:, 1:] // 2], 1), dim=0)
        y = Sp(None, coarse, skip=x)
        return y.replace_feature(y.read(out))


class SparseInverseConv3d(_Conv):
    def forward(self, x):
        fine = x.skip
        out = F.conv_transpose3d(x.dense(), self.weight.permute(4, 0, 1, 2, 3), self.bias, stride=2)
        return Sp(fine.read(out), fine.coords, fine.skip)


class BasicBlock(nn.Module):
    def __init__(self, in_channels, embed_channels, norm_fn=None, **_):
        super().__init__()
        self.in_channels, self.embed_channels = in_channels, embed_channels
        if in_channels == embed_channels:
            self.proj = nn.Sequential(nn.Identity())
        else:
            self.proj_conv = SubMConv3d(in_channels, embed_channels, 1)
            self.proj_norm = norm_fn(embed_channels)
        self.conv1 = SubMConv3d(in_channels, embed_channels, 3)
        self.bn1 = norm_fn(embed_channels)
        self.relu = nn.ReLU()
        self.conv2 = SubMConv3d(embed_channels, embed_channels, 3)
        self.bn2 = norm_fn(embed_channels)

    def forward(self, x):
        x, condition, context = x
        residual = x
        out = self.conv1(x)
        out = out.replace_feature(self.relu(self.bn1(out.features, condition, context)))
        out = self.conv2(out)
        out = out.replace_feature(self.bn2(out.features, condition, context))
        if self.in_channels != self.embed_channels:
            residual = residual.replace_feature(self.proj_norm(self.proj_conv(residual).features, condition, context))
        return out.replace_feature(self.relu(out.features + residual.features)), condition, context


class _ConvBnRelu(nn.Module):
    def __init__(self, conv, cout, norm_fn):
        super().__init__()
        self.conv, self.bn, self.relu = conv, norm_fn(cout), nn.ReLU()

    def forward(self, x):
        x, condition, context = x
        out = self.conv(x)
        return out.replace_feature(self.relu(self.bn(out.features, condition, context)))


class _Blocks(nn.Module):
    def __init__(self, blocks):
        super().__init__()
        for k, v in blocks.items():
            self.add_module(k, v)

    def forward(self, x):
        for m in self.children():
            x = m(x)
        return x


class OracleSpUNet(nn.Module):
    def __init__(self, in_channels, num_classes=0, base_channels=32, context_channels=256,
                 channels=(32, 64, 128, 256, 256, 128, 96, 96), layers=(2, 3, 4, 6, 2, 2, 2, 2), cls_mode=False,
                 conditions=("ScanNet", "S3DIS", "Structured3D"), zero_init=False, norm_decouple=True, norm_adaptive=True,
                 norm_affine=True, pretrained_path=None):
        super().__init__()
        self.num_stages, self.cls_mode, self.conditions, self.num_classes = len(layers) // 2, cls_mode, conditions, num_classes
        self.embedding_table = nn.Embedding(len(conditions), context_channels) if norm_adaptive else None
        norm_fn = partial(PDBatchNorm, eps=1e-3, momentum=0.01, conditions=conditions, context_channels=context_channels,
                          decouple=norm_decouple, adaptive=norm_adaptive, affine=norm_affine)
        self.conv_input = _ConvBnRelu(SubMConv3d(in_channels, base_channels, 5), base_channels, norm_fn)
        enc_c, dec_c = base_channels, channels[-1]
        self.down, self.up, self.enc = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.dec = nn.ModuleList() if not cls_mode else None
        for s in range(self.num_stages):
            self.down.append(_ConvBnRelu(SparseConv3d(enc_c, channels[s], 2), channels[s], norm_fn))
            self.enc.append(_Blocks(OrderedDict((f"block{i}", BasicBlock(channels[s], channels[s], norm_fn=norm_fn)) for i in range(layers[s]))))
            if not cls_mode:
                self.up.append(_ConvBnRelu(SparseInverseConv3d(channels[len(channels) - s - 2], dec_c, 2), dec_c, norm_fn))
                self.dec.append(_Blocks(OrderedDict((f"block{i}", BasicBlock(dec_c + enc_c if i == 0 else dec_c, dec_c, norm_fn=norm_fn))
                                                    for i in range(layers[len(channels) - s - 1]))))
            enc_c, dec_c = channels[s], channels[len(channels) - s - 2]
        final_in = channels[-1] if not cls_mode else channels[self.num_stages - 1]
        self.final = SubMConv3d(final_in, num_classes, 1, bias=True) if num_classes > 0 else nn.Identity()
        self.num_channels = num_classes if num_classes > 0 else final_in

    def forward(self, input_dict):
        grid_coord, feat, offset = input_dict["grid_coord"], input_dict["feat"], input_dict["offset"]
        condition = input_dict["condition"][0] if "condition" in input_dict else self.conditions[0]
        context = input_dict.get("context", None)
        if context is None and self.embedding_table is not None:
            context = self.embedding_table(torch.tensor([self.conditions.index(condition)]))
        sizes = torch.diff(offset, prepend=offset.new_zeros(1))
        batch = torch.repeat_interleave(torch.arange(len(sizes)), sizes)
        x = Sp(feat, torch.cat([batch[:, None], grid_coord], 1).long())
        x = self.conv_input([x, condition, context])
        skips = [x]
        for s in range(self.num_stages):
            x = self.down[s]([x, condition, context])
            x, _, _ = self.enc[s]([x, condition, context])
            skips.append(x)
        x = skips.pop(-1)
        if not self.cls_mode:
            for s in reversed(range(self.num_stages)):
                x = self.up[s]([x, condition, context])
                skip = skips.pop(-1)
                x = x.replace_feature(torch.cat((x.features, skip.features), dim=1))
                x, _, _ = self.dec[s]([x, condition, context])
        if self.num_classes > 0:
            x = self.final(x)
        f = x.features
        if self.cls_mode:
            b = x.coords[:, 0]
            nb = len(sizes)
            f = torch.zeros(nb, f.shape[1]).index_add_(0, b, f) / torch.bincount(b, minlength=nb).clamp_min(1)[:, None]
        return f
