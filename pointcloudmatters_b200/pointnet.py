"""PointNet backbone -- host-side mirror of `src/models/components/pcd_encoder/pointnet.py:16-85`.

The reference expresses this per-point MLP (6->64->64->64->128->512, each layer
`SubMConv3d(kernel_size=1, bias=False)` + `BatchNorm1d(eps=1e-3, momentum=0.01)` + ReLU) through
spconv; with unique voxels a k=1 submanifold convolution is a row-wise Linear, so no rule
generation / hashing is needed at all.  Parameter names and shapes follow the reference
(`conv1.0.weight` in spconv's (out, 1, 1, 1, in) layout, `conv1.1.*` BatchNorm), so its
checkpoints load unchanged.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as PF


class _Conv(nn.Module):
    def __init__(self, cin, cout, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, 1, 1, 1, cin))
        nn.init.kaiming_uniform_(self.weight.view(cout, cin), a=math.sqrt(5))
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # accept spconv-1.x (k, k, k, in, out) layout as well as 2.x (out, k, k, k, in)
        key = prefix + "weight"
        if key in state_dict and tuple(state_dict[key].shape) == (1, 1, 1, self.weight.shape[-1], self.weight.shape[0]):
            state_dict[key] = state_dict[key].permute(4, 0, 1, 2, 3).contiguous()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    @property
    def matrix(self):
        return self.weight.view(self.weight.shape[0], -1)


class PointNet(nn.Module):
    def __init__(self, in_channels, num_classes=0, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.embedding_table = None
        dims = [in_channels, 64, 64, 64, 128, 512]
        for i in range(5):
            setattr(self, f"conv{i + 1}", nn.Sequential(_Conv(dims[i], dims[i + 1]),
                                                        nn.BatchNorm1d(dims[i + 1], eps=1e-3, momentum=0.01)))
        self.final = _Conv(512, num_classes, bias=True) if num_classes > 0 else nn.Identity()
        self.num_channels = num_classes if num_classes > 0 else 512

    def forward(self, input_dict):
        x = input_dict["feat"]
        for i in range(5):
            conv, bn = getattr(self, f"conv{i + 1}")
            y = PF.linear(x, conv.matrix)
            if bn.training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            x = PF.batchnorm_relu(y, bn)
        if self.num_classes > 0:
            x = PF.linear(x, self.final.matrix, self.final.bias)
        return x
