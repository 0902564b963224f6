"""SpUNet sparse-convolution encoder -- host-side mirror of `src/models/components/pcd_encoder/spunet.py:19-463`
(`PDBatchNorm`, `BasicBlock`, `SPConvDown`, `SPConvUp`, `SPConvPatchEmbedding`, `SpUNet`): same constructor kwargs, same
module tree and `state_dict` keys / shapes (spconv 2.x weight layout `(out, kD, kH, kW, in)`), same
`forward(input_dict) -> (N, num_channels)` contract, so `_target_: ...spunet.SpUNet` configs and PonderV2 checkpoints
(`load_ponderv2_weights`, :399-409) carry over.

The reference runs on the third-party `spconv` library (not vendored, not pinned, absent from this image).  Here every
sparse convolution is  rules (csrc/spconv.cu: voxel hash table, submanifold neighbour table, stride-2 parent / child
tables)  ->  gather into a bf16 column matrix  ->  ONE tcgen05 GEMM against the weight read in place  (backward: two GEMMs
+ an atomic-free transpose gather).  PDBatchNorm (+ its FiLM modulation) and ReLU run as the fused BatchNorm kernels of
csrc/batchnorm.cu with effective per-channel (gamma, beta).

PARITY UNPINNED: spconv is absent, so this module is checked against a dense-voxel restatement (oracle/spunet_oracle.py:
F.conv3d / conv_transpose3d on the densified grid), not against spconv itself.  Assumptions about spconv that the
oracle shares: SubMConv3d is centred for any `padding` argument (output set = input set, offsets -k/2 .. k/2);
SparseConv3d(k=2, s=2) maps voxel c to c // 2 with kernel offset c % 2; SparseInverseConv3d reuses those pairs.
Data-dependent level sizes cost one device->host read per resolution level, so a SpUNet policy runs the eager step
(no CUDA-graph capture).
"""
from __future__ import annotations

from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn

from . import functional as PF
from . import kernels as K
from ._lib import PcmError, check, current_stream, lib, ptr

_INT_MAX = 2 ** 31 - 1


# ----------------------------------------------------------------------------------------------------------------
# rules
# ----------------------------------------------------------------------------------------------------------------
class SparseLevels:
    """Voxel sets of the U-Net's resolution levels plus the rule tables derived from them (built lazily, cached per
    (level, kind) like spconv's `indice_key`)."""

    def __init__(self, coords):
        self.coords = [coords.contiguous()]  # level 0: (n, 4) int32 [batch, x, y, z]
        self._tables, self._subm, self._down = {}, {}, {}

    @staticmethod
    def _cap(n):
        c = 16
        while c < 2 * n:
            c *= 2
        return c

    def table(self, level, shift=0):
        key = (level, shift)
        if key not in self._tables:
            c = self.coords[level]
            n, dev = c.shape[0], c.device
            cap = self._cap(n)
            tkey = torch.full((cap,), -1, dtype=torch.int64, device=dev)
            tval = torch.full((cap,), _INT_MAX, dtype=torch.int32, device=dev)
            check(lib.pcm_spconv_build_table(n, ptr(c), shift, ptr(tkey), ptr(tval), cap, current_stream()), "pcm_spconv_build_table")
            self._tables[key] = (tkey, tval, cap)
        return self._tables[key]

    def subm(self, level, k):
        """nbr (n, k^3) int32 of the submanifold convolution with kernel size k on `level`."""
        key = (level, k)
        if key not in self._subm:
            c = self.coords[level]
            n = c.shape[0]
            tkey, tval, cap = self.table(level)
            nbr = torch.empty((n, k ** 3), dtype=torch.int32, device=c.device)
            check(lib.pcm_spconv_subm_rules(n, k, ptr(c), ptr(tkey), ptr(tval), cap, ptr(nbr), current_stream()), "pcm_spconv_subm_rules")
            self._subm[key] = nbr
        return self._subm[key]

    def down(self, level):
        """(parent (n), kidx (n), child (m, 8)) of the stride-2 convolution level -> level + 1; creates level + 1."""
        if level not in self._down:
            c = self.coords[level]
            n, dev = c.shape[0], c.device
            tkey, tval, cap = self.table(level, shift=1)
            i32 = lambda *s: torch.empty(s, dtype=torch.int32, device=dev)
            leader, excl, parent, kidx, coarse, m_out = i32(n), i32(n), i32(n), i32(n), i32(n, 4), i32(1)
            child = torch.full((n, 8), -1, dtype=torch.int32, device=dev)
            check(lib.pcm_spconv_down_rules(n, ptr(c), ptr(tkey), ptr(tval), cap, ptr(leader), ptr(excl), ptr(parent), ptr(kidx),
                                            ptr(child), ptr(coarse), ptr(m_out), current_stream()), "pcm_spconv_down_rules")
            m = int(m_out.item())  # the level's size is data dependent: one device->host read per level
            self._down[level] = (parent, kidx, child[:m].contiguous())
            if len(self.coords) == level + 1:
                self.coords.append(coarse[:m].contiguous())
        return self._down[level]


class SparseTensor:
    """(features, level) pair standing in for spconv.SparseConvTensor."""

    def __init__(self, features, levels, level):
        self.features, self.levels, self.level = features, levels, level

    def replace_feature(self, f):
        return SparseTensor(f, self.levels, self.level)


# ----------------------------------------------------------------------------------------------------------------
# sparse convolution = gather + tcgen05 GEMM
# ----------------------------------------------------------------------------------------------------------------
class _SparseConv(torch.autograd.Function):
    """mode 'subm' / 'down': y[i] = sum_o W[:, o, :] x[tab[i, o]] (+ bias);  mode 'inverse': y[i] = W[:, kidx[i], :] x[parent[i]]."""

    @staticmethod
    def forward(ctx, x, weight, bias, mode, tab, parent, kidx, n_out):
        Cout, Cin = weight.shape[0], weight.shape[-1]
        kvol = weight.numel() // (Cout * Cin)
        dev = x.device
        xc = x.contiguous()
        wb = PF._wb(weight)
        ctx.mode, ctx.dims, ctx.params = mode, (Cout, Cin, kvol, n_out, x.shape[0]), (weight, bias)
        if mode == "inverse":
            xb = K.add_cast_bf16(xc)
            Z = K.gemm_bf16(xb, wb.view(Cout * kvol, Cin))  # (M, Cout * 8) fp32, column = co * 8 + kk
            out = torch.empty((n_out, Cout), dtype=torch.float32, device=dev)
            check(lib.pcm_spconv_inverse_pick(n_out, Cout, ptr(Z), ptr(parent), ptr(kidx), ptr(out), current_stream()),
                  "pcm_spconv_inverse_pick")
            if bias is not None:
                out = out + bias
            ctx.save_for_backward(xb, wb, tab)
            return out
        Cp = -(-Cin // 8) * 8
        col = torch.empty((n_out, kvol * Cp), dtype=torch.bfloat16, device=dev)
        check(lib.pcm_spconv_gather(n_out, kvol, Cin, Cp, ptr(xc), xc.stride(0), int(xc.dtype == torch.bfloat16), ptr(tab), ptr(col),
                                    current_stream()), "pcm_spconv_gather")
        wm = wb.view(Cout, kvol * Cin)
        if Cp != Cin:  # narrow input (the stem's 6 channels): zero-padded weight copy matching the padded columns
            wm = torch.nn.functional.pad(wb.view(Cout, kvol, Cin), (0, Cp - Cin)).reshape(Cout, kvol * Cp)
        y = K.gemm_bf16(col, wm, bias=bias)
        ctx.save_for_backward(col, wm, tab, parent, kidx)
        ctx.Cp = Cp
        return y

    @staticmethod
    def backward(ctx, dy):
        Cout, Cin, kvol, n_out, n_in = ctx.dims
        weight, bias = ctx.params
        dev = dy.device
        dyc = dy.contiguous()
        db = dyc.sum(0) if (bias is not None and ctx.needs_input_grad[2]) else None
        slot = PF._grad_slot(weight)
        if ctx.mode == "inverse":
            xb, wb, child = ctx.saved_tensors
            M = xb.shape[0]
            dZ = torch.empty((M, Cout * kvol), dtype=torch.bfloat16, device=dev)
            check(lib.pcm_spconv_inverse_place(M, Cout, ptr(dyc), ptr(child), ptr(dZ), current_stream()), "pcm_spconv_inverse_place")
            wm = wb.view(Cout * kvol, Cin)
            dx = K.gemm_bf16(dZ, wm, b_mn=True) if ctx.needs_input_grad[0] else None
            dW = slot.view(Cout * kvol, Cin) if slot is not None else torch.zeros((Cout * kvol, Cin), dtype=torch.float32, device=dev)
            K.gemm_bf16(dZ, xb, a_mn=True, b_mn=True, out=dW, accumulate=True, split_k=0)
            return dx, (None if slot is not None else dW.view(weight.shape)), db, None, None, None, None, None
        col, wm, tab, parent, kidx = ctx.saved_tensors
        Cp = ctx.Cp
        dyb = K.add_cast_bf16(dyc)
        dx = None
        if ctx.needs_input_grad[0]:
            dcol = K.gemm_bf16(dyb, wm, b_mn=True)  # (n_out, kvol * Cp) fp32
            dx = torch.empty((n_in, Cin), dtype=torch.float32, device=dev)
            mode = 0 if ctx.mode == "subm" else 1
            check(lib.pcm_spconv_gather_bwd(n_in, kvol, Cin, Cp, mode, ptr(dcol), ptr(tab), ptr(parent), ptr(kidx), ptr(dx),
                                            current_stream()), "pcm_spconv_gather_bwd")
        if Cp == Cin and slot is not None:
            K.gemm_bf16(dyb, col, a_mn=True, b_mn=True, out=slot.view(Cout, kvol * Cin), accumulate=True, split_k=0)
            dW = None
        else:
            dWp = torch.zeros((Cout, kvol * Cp), dtype=torch.float32, device=dev)
            K.gemm_bf16(dyb, col, a_mn=True, b_mn=True, out=dWp, accumulate=True, split_k=0)
            dW = dWp.view(Cout, kvol, Cp)[:, :, :Cin].reshape(weight.shape)
            if slot is not None:
                slot.add_(dW)
                dW = None
        return dx, dW, db, None, None, None, None, None


class _SparseConvBase(nn.Module):
    """Parameter container with spconv 2.x's layout: weight (out, k, k, k, in), optional bias (out)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=False, indice_key=None, **_ignored):
        super().__init__()
        k = kernel_size
        self.in_channels, self.out_channels, self.kernel_size, self.indice_key = in_channels, out_channels, k, indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))
        nn.init.trunc_normal_(self.weight, std=0.02)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "weight"
        k, ci, co = self.kernel_size, self.in_channels, self.out_channels
        if key in state_dict and tuple(state_dict[key].shape) == (k, k, k, ci, co) and (k, k, k, ci, co) != (co, k, k, k, ci):
            state_dict[key] = state_dict[key].permute(4, 0, 1, 2, 3).contiguous()  # spconv 1.x layout
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class SubMConv3d(_SparseConvBase):
    def forward(self, x):
        if self.kernel_size == 1:
            y = PF.linear(x.features, self.weight.view(self.out_channels, self.in_channels), self.bias)
            return x.replace_feature(y)
        nbr = x.levels.subm(x.level, self.kernel_size)
        n = x.features.shape[0]
        return x.replace_feature(_SparseConv.apply(x.features, self.weight, self.bias, "subm", nbr, None, None, n))


class SparseConv3d(_SparseConvBase):
    def forward(self, x):
        if self.kernel_size != 2:
            raise NotImplementedError("SpUNet downsamples with kernel_size = stride = 2 only (spunet.py:159-166)")
        parent, kidx, child = x.levels.down(x.level)
        y = _SparseConv.apply(x.features, self.weight, self.bias, "down", child, parent, kidx, child.shape[0])
        return SparseTensor(y, x.levels, x.level + 1)


class SparseInverseConv3d(_SparseConvBase):
    def forward(self, x):
        parent, kidx, child = x.levels.down(x.level - 1)
        y = _SparseConv.apply(x.features, self.weight, self.bias, "inverse", child, parent, kidx, parent.shape[0])
        return SparseTensor(y, x.levels, x.level - 1)


# ----------------------------------------------------------------------------------------------------------------
# modules of spunet.py
# ----------------------------------------------------------------------------------------------------------------
class _TouchParams(torch.autograd.Function):
    """Identity on `x` that makes `params` part of the graph with an exactly-zero gradient: the reference multiplies the
    non-selected conditions' BatchNorm outputs by 0 (spunet.py:58-62), so their affine parameters receive zero gradients
    (and are therefore weight-decayed by AdamW) instead of none."""

    @staticmethod
    def forward(ctx, x, *params):
        ctx.shapes = [p.shape for p in params]
        ctx.meta = (x.dtype, x.device)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        dt, dev = ctx.meta
        return (g, *[torch.zeros(s, dtype=dt, device=dev) for s in ctx.shapes])


class PDBatchNorm(nn.Module):
    """spunet.py:19-73: per-condition BatchNorm1d copies (all of them see the input in training mode, the selected one
    produces the output) + optional FiLM modulation from a 256-d context.  `relu=True` fuses the following ReLU."""

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

    def forward(self, feat, condition=None, context=None, relu=False):
        if self.decouple:
            assert condition in self.conditions
            sel = self.conditions.index(condition)
            bn = self.bns[sel]
            others = [b for i, b in enumerate(self.bns) if i != sel]
        else:
            bn, others = self.bn, []
        C = feat.shape[1]
        training = bn.training or bn.running_mean is None
        gamma = bn.weight if bn.affine else torch.ones(C, dtype=torch.float32, device=feat.device)
        beta = bn.bias if bn.affine else torch.zeros(C, dtype=torch.float32, device=feat.device)
        touched = [t for b in others if b.affine for t in (b.weight, b.bias)]
        if touched and torch.is_grad_enabled():
            gamma = _TouchParams.apply(gamma, *touched)
        if self.adaptive:
            assert context is not None
            shift, scale = self.modulation(context).chunk(2, dim=1)  # (1, C) each: feat * (1 + scale) + shift
            gamma = gamma * (1.0 + scale[0])
            beta = beta * (1.0 + scale[0]) + shift[0]
        extra = None
        if training:
            for b in [bn] + others:
                if b.num_batches_tracked is not None:
                    b.num_batches_tracked.add_(1)
            extra = [(b.running_mean, b.running_var, b.momentum if b.momentum is not None else 0.0) for b in others
                     if b.running_mean is not None]
        if not (feat.is_cuda and feat.dtype == torch.float32 and C % 4 == 0 and C <= 1024):
            raise PcmError("PDBatchNorm runs on CUDA fp32 features with C % 4 == 0 (no CPU fallback)")
        out, outb = PF._BatchNormReLU.apply(feat.contiguous(), gamma.contiguous(), beta.contiguous(), bn.running_mean, bn.running_var,
                                            bn.eps, bn.momentum if bn.momentum is not None else 0.0, training, relu, extra)
        out._pcm_bf16 = outb
        return out


class BasicBlock(nn.Module):
    """spunet.py:76-146."""
    expansion = 1

    def __init__(self, in_channels, embed_channels, stride=1, norm_fn=None, indice_key=None, bias=False):
        super().__init__()
        assert norm_fn is not None
        self.in_channels, self.embed_channels = in_channels, embed_channels
        if in_channels == embed_channels:
            self.proj = nn.Sequential(nn.Identity())
        else:
            self.proj_conv = SubMConv3d(in_channels, embed_channels, kernel_size=1, bias=False)
            self.proj_norm = norm_fn(embed_channels)
        self.conv1 = SubMConv3d(in_channels, embed_channels, kernel_size=3, bias=bias, indice_key=indice_key)
        self.bn1 = norm_fn(embed_channels)
        self.relu = nn.ReLU()
        self.conv2 = SubMConv3d(embed_channels, embed_channels, kernel_size=3, bias=bias, indice_key=indice_key)
        self.bn2 = norm_fn(embed_channels)
        self.stride = stride

    def forward(self, x):
        x, condition, context = x
        residual = x.features
        out = self.conv1(x)
        out = out.replace_feature(self.bn1(out.features, condition, context, relu=True))
        out = self.conv2(out)
        f = self.bn2(out.features, condition, context)
        if self.in_channels != self.embed_channels:
            residual = self.proj_norm(self.proj_conv(x).features, condition, context)
        return out.replace_feature(torch.relu(f + residual)), condition, context


class SPConvDown(nn.Module):
    """spunet.py:149-175."""

    def __init__(self, in_channels, out_channels, indice_key, kernel_size=2, bias=False, norm_fn=None):
        super().__init__()
        self.conv = SparseConv3d(in_channels, out_channels, kernel_size=kernel_size, bias=bias, indice_key=indice_key)
        self.bn = norm_fn(out_channels)
        self.relu = nn.ReLU()

    def forward(self, x):
        x, condition, context = x
        out = self.conv(x)
        return out.replace_feature(self.bn(out.features, condition, context, relu=True))


class SPConvUp(nn.Module):
    """spunet.py:178-203."""

    def __init__(self, in_channels, out_channels, indice_key, kernel_size=2, bias=False, norm_fn=None):
        super().__init__()
        self.conv = SparseInverseConv3d(in_channels, out_channels, kernel_size=kernel_size, bias=bias, indice_key=indice_key)
        self.bn = norm_fn(out_channels)
        self.relu = nn.ReLU()

    def forward(self, x):
        x, condition, context = x
        out = self.conv(x)
        return out.replace_feature(self.bn(out.features, condition, context, relu=True))


class SPConvPatchEmbedding(nn.Module):
    """spunet.py:206-226."""

    def __init__(self, in_channels, out_channels, kernel_size=5, norm_fn=None):
        super().__init__()
        self.conv = SubMConv3d(in_channels, out_channels, kernel_size=kernel_size, bias=False, indice_key="stem")
        self.bn = norm_fn(out_channels)
        self.relu = nn.ReLU()

    def forward(self, x):
        x, condition, context = x
        out = self.conv(x)
        return out.replace_feature(self.bn(out.features, condition, context, relu=True))


class _Blocks(nn.Module):
    """spconv.SparseSequential(OrderedDict(block0=..., block1=...)): same child names."""

    def __init__(self, blocks: OrderedDict):
        super().__init__()
        for k, v in blocks.items():
            self.add_module(k, v)

    def forward(self, x):
        for m in self.children():
            x = m(x)
        return x


class SpUNet(nn.Module):
    """spunet.py:229-463."""

    def __init__(self, in_channels, num_classes=0, base_channels=32, context_channels=256,
                 channels=(32, 64, 128, 256, 256, 128, 96, 96), layers=(2, 3, 4, 6, 2, 2, 2, 2), cls_mode=False,
                 conditions=("ScanNet", "S3DIS", "Structured3D"), zero_init=False, norm_decouple=True, norm_adaptive=True,
                 norm_affine=True, pretrained_path=None):
        super().__init__()
        assert len(layers) % 2 == 0 and len(layers) == len(channels)
        self.in_channels, self.num_classes, self.base_channels = in_channels, num_classes, base_channels
        self.channels, self.layers, self.num_stages = channels, layers, len(layers) // 2
        self.cls_mode, self.conditions, self.zero_init = cls_mode, conditions, zero_init
        self.embedding_table = nn.Embedding(len(conditions), context_channels) if norm_adaptive else None
        norm_fn = partial(PDBatchNorm, eps=1e-3, momentum=0.01, conditions=conditions, context_channels=context_channels,
                          decouple=norm_decouple, adaptive=norm_adaptive, affine=norm_affine)
        self.conv_input = SPConvPatchEmbedding(in_channels, base_channels, kernel_size=5, norm_fn=norm_fn)
        enc_channels, dec_channels = base_channels, channels[-1]
        self.down, self.up, self.enc = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.dec = nn.ModuleList() if not cls_mode else None
        for s in range(self.num_stages):
            self.down.append(SPConvDown(enc_channels, channels[s], kernel_size=2, bias=False, indice_key=f"spconv{s + 1}", norm_fn=norm_fn))
            self.enc.append(_Blocks(OrderedDict((f"block{i}", BasicBlock(channels[s], channels[s], norm_fn=norm_fn, indice_key=f"subm{s + 1}"))
                                                for i in range(layers[s]))))
            if not cls_mode:
                self.up.append(SPConvUp(channels[len(channels) - s - 2], dec_channels, kernel_size=2, bias=False,
                                        indice_key=f"spconv{s + 1}", norm_fn=norm_fn))
                self.dec.append(_Blocks(OrderedDict(
                    (f"block{i}", BasicBlock(dec_channels + enc_channels if i == 0 else dec_channels, dec_channels, norm_fn=norm_fn,
                                             indice_key=f"subm{s}")) for i in range(layers[len(channels) - s - 1]))))
            enc_channels, dec_channels = channels[s], channels[len(channels) - s - 2]
        final_in = channels[-1] if not cls_mode else channels[self.num_stages - 1]
        self.final = SubMConv3d(final_in, num_classes, kernel_size=1, bias=True) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)
        self.num_channels = num_classes if num_classes > 0 else final_in
        if pretrained_path is not None:
            self.load_ponderv2_weights(pretrained_path)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, SubMConv3d):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm1d):
            if m.affine:
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        elif isinstance(m, PDBatchNorm):
            if self.zero_init:
                nn.init.constant_(m.modulation[-1].weight, 0)
                nn.init.constant_(m.modulation[-1].bias, 0)

    def load_ponderv2_weights(self, path):
        """spunet.py:399-409."""
        weight = OrderedDict()
        checkpoint = torch.load(path, map_location="cpu")
        for key, value in checkpoint["state_dict"].items():
            if key.startswith("module.backbone."):
                weight[key.replace("module.backbone.", "")] = value
            elif key.startswith("module.embedding_table"):
                weight[key.replace("module.", "")] = value
        self.load_state_dict(weight, strict=True)

    def forward(self, input_dict):
        grid_coord, feat, offset = input_dict["grid_coord"], input_dict["feat"], input_dict["offset"]
        if not feat.is_cuda:
            raise PcmError("SpUNet runs on CUDA tensors only (no CPU fallback)")
        condition = input_dict["condition"][0] if "condition" in input_dict else self.conditions[0]
        if "context" in input_dict:
            context = input_dict["context"]
        elif self.embedding_table is not None:
            context = self.embedding_table(torch.tensor([self.conditions.index(condition)], device=grid_coord.device))
        else:
            context = None
        n = feat.shape[0]
        batch = torch.searchsorted(offset.to(torch.int64), torch.arange(n, device=feat.device), right=True)  # offset2batch, sync-free
        coords = torch.cat([batch.unsqueeze(-1).int(), grid_coord.int()], dim=1).contiguous()
        x = SparseTensor(feat.float(), SparseLevels(coords), 0)
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
        feats = x.features
        if self.cls_mode:  # scatter(mean) over the batch index (torch_geometric.utils.scatter, spunet.py:459-462)
            b_idx = x.levels.coords[x.level][:, 0].long()
            nb = int(offset.shape[0])
            summed = torch.zeros((nb, feats.shape[1]), dtype=feats.dtype, device=feats.device).index_add_(0, b_idx, feats)
            cnt = torch.zeros(nb, dtype=feats.dtype, device=feats.device).index_add_(0, b_idx, torch.ones_like(b_idx, dtype=feats.dtype))
            feats = summed / cnt.clamp_min(1).unsqueeze(1)
        return feats
