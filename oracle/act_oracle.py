"""oracle/act_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain-PyTorch (explicit-math, device-agnostic, runs on CPU) restatement of the reference's
behaviour-cloning policy for the point-cloud modality:

  * `OraclePointNet`      <- src/models/components/pcd_encoder/pointnet.py:16-85.  The reference
                             builds it from spconv `SubMConv3d(kernel_size=1, bias=False)` layers
                             (spconv is an un-vendored, unpinned dependency, README.md:119-123);
                             with unique voxel coordinates per cloud a k=1 submanifold conv is a
                             row-wise Linear, which is what is restated here.
  * `OracleTransformer*`  <- src/models/components/act/transformer.py:16-425 (DETR-style post-LN
                             encoder / decoder; nn.MultiheadAttention restated as explicit
                             softmax(QK^T/sqrt(d) + mask) V).
  * `OracleACTPCD`,
    `OracleACTRLBenchPCD` <- src/models/components/act/act.py:40-309, :312-598, :707-825.

Parameter / buffer names and shapes equal the reference's `state_dict` so the same checkpoint
loads into reference, oracle and product modules.  The point-set operators (FPS, kNN) come from
the C oracle (oracle/pointops_oracle.c) -- never from the CUDA product path.

Parity pin: tests/golden/act_*.npz hold inputs, a state_dict and outputs produced by the
REFERENCE modules themselves (imported from /root/reference by oracle/gen_golden_act.py in the
build container); tests/test_act_oracle_cpu.py checks this restatement against them.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointops_oracle as PO


# ----------------------------------------------------------------------------------------------
# PointNet (per-point MLP)
# ----------------------------------------------------------------------------------------------
class _ConvBNReLU(nn.Module):
    """Index 0 holds the conv weight in spconv's (out, 1, 1, 1, in) layout, index 1 the BN."""

    def __init__(self, cin, cout):
        super().__init__()
        conv = nn.Module()
        conv.weight = nn.Parameter(torch.empty(cout, 1, 1, 1, cin))
        nn.init.kaiming_uniform_(conv.weight.view(cout, cin), a=math.sqrt(5))
        self.add_module("0", conv)
        self.add_module("1", nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01))  # pointnet.py:29

    def forward(self, x):
        w = getattr(self, "0").weight
        return F.relu(getattr(self, "1")(x @ w.view(w.shape[0], -1).t()))


class OraclePointNet(nn.Module):
    def __init__(self, in_channels, num_classes=0, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        dims = [in_channels, 64, 64, 64, 128, 512]  # pointnet.py:31-55
        for i in range(5):
            setattr(self, f"conv{i + 1}", _ConvBNReLU(dims[i], dims[i + 1]))
        if num_classes > 0:  # pointnet.py:57-61 (bias=True)
            self.final = nn.Module()
            self.final.weight = nn.Parameter(torch.empty(num_classes, 1, 1, 1, 512))
            self.final.bias = nn.Parameter(torch.zeros(num_classes))
            nn.init.kaiming_uniform_(self.final.weight.view(num_classes, 512), a=math.sqrt(5))
        self.num_channels = num_classes if num_classes > 0 else 512

    def forward(self, input_dict):
        x = input_dict["feat"]
        for i in range(5):
            x = getattr(self, f"conv{i + 1}")(x)
        if self.num_classes > 0:
            x = x @ self.final.weight.view(self.num_classes, -1).t() + self.final.bias
        return x


# ----------------------------------------------------------------------------------------------
# Transformer
# ----------------------------------------------------------------------------------------------
class _MHAParams(nn.Module):
    """Same parameter names as nn.MultiheadAttention (in_proj_weight/bias, out_proj.weight/bias)."""

    def __init__(self, d_model, nhead, dropout):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout = d_model, nhead, dropout
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, query, key, value, key_padding_mask=None):
        L, B, E = query.shape
        S = key.shape[0]
        h, d = self.num_heads, E // self.num_heads
        wq, wk, wv = self.in_proj_weight.split(E, 0)
        bq, bk, bv = self.in_proj_bias.split(E, 0)
        q = (query @ wq.t() + bq).view(L, B, h, d).permute(1, 2, 0, 3)
        k = (key @ wk.t() + bk).view(S, B, h, d).permute(1, 2, 0, 3)
        v = (value @ wv.t() + bv).view(S, B, h, d).permute(1, 2, 0, 3)
        scores = (q * (1.0 / math.sqrt(d))) @ k.transpose(-1, -2)  # torch scales q first
        if key_padding_mask is not None:
            scores = scores.masked_fill(key_padding_mask.view(B, 1, 1, S), float("-inf"))
        attn = F.dropout(torch.softmax(scores, dim=-1), self.dropout, self.training)
        out = (attn @ v).permute(2, 0, 1, 3).reshape(L, B, E)
        return self.out_proj(out)


def _act(name):
    return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]


class OracleEncoderLayer(nn.Module):
    """transformer.py:210-283."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = _MHAParams(d_model, nhead, dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1, self.norm2 = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.p, self.activation, self.normalize_before = dropout, _act(activation), normalize_before

    def _drop(self, x):
        return F.dropout(x, self.p, self.training)

    def forward(self, src, src_key_padding_mask=None, pos=None):
        add = (lambda t: t if pos is None else t + pos)
        if self.normalize_before:  # forward_pre :256-270
            s2 = self.norm1(src)
            qk = add(s2)
            src = src + self._drop(self.self_attn(qk, qk, s2, src_key_padding_mask))
            s2 = self.norm2(src)
            return src + self._drop(self.linear2(self._drop(self.activation(self.linear1(s2)))))
        qk = add(src)  # forward_post :238-254
        src = self.norm1(src + self._drop(self.self_attn(qk, qk, src, src_key_padding_mask)))
        return self.norm2(src + self._drop(self.linear2(self._drop(self.activation(self.linear1(src))))))


class OracleTransformerEncoder(nn.Module):
    """transformer.py:118-158."""

    def __init__(self, d_model=256, nhead=8, dim_feedforward=2048, dropout=0.1, activation="relu",
                 normalize_before=False, num_layers=4):
        super().__init__()
        self.layers = nn.ModuleList([OracleEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation,
                                                        normalize_before) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = nn.LayerNorm(d_model) if normalize_before else None

    def forward(self, src, mask=None, src_key_padding_mask=None, pos=None):
        out = src
        for layer in self.layers:
            out = layer(out, src_key_padding_mask=src_key_padding_mask, pos=pos)
        return out if self.norm is None else self.norm(out)


class OracleDecoderLayer(nn.Module):
    """transformer.py:286-404."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = _MHAParams(d_model, nhead, dropout)
        self.multihead_attn = _MHAParams(d_model, nhead, dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d_model), nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.p, self.activation, self.normalize_before = dropout, _act(activation), normalize_before

    def _drop(self, x):
        return F.dropout(x, self.p, self.training)

    def forward(self, tgt, memory, memory_key_padding_mask=None, pos=None, query_pos=None):
        wq = (lambda t: t if query_pos is None else t + query_pos)
        wp = (lambda t: t if pos is None else t + pos)
        if self.normalize_before:  # forward_pre :348-375
            t2 = self.norm1(tgt)
            qk = wq(t2)
            tgt = tgt + self._drop(self.self_attn(qk, qk, t2))
            t2 = self.norm2(tgt)
            tgt = tgt + self._drop(self.multihead_attn(wq(t2), wp(memory), memory, memory_key_padding_mask))
            t2 = self.norm3(tgt)
            return tgt + self._drop(self.linear2(self._drop(self.activation(self.linear1(t2)))))
        qk = wq(tgt)  # forward_post :317-346
        tgt = self.norm1(tgt + self._drop(self.self_attn(qk, qk, tgt)))
        tgt = self.norm2(tgt + self._drop(self.multihead_attn(wq(tgt), wp(memory), memory, memory_key_padding_mask)))
        return self.norm3(tgt + self._drop(self.linear2(self._drop(self.activation(self.linear1(tgt))))))


class OracleTransformerDecoder(nn.Module):
    """transformer.py:161-207."""

    def __init__(self, d_model, nhead, dim_feedforward, dropout, activation, normalize_before, num_layers,
                 return_intermediate):
        super().__init__()
        self.layers = nn.ModuleList([OracleDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation,
                                                        normalize_before) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(d_model)
        self.return_intermediate = return_intermediate

    def forward(self, tgt, memory, memory_key_padding_mask=None, pos=None, query_pos=None):
        out, inter = tgt, []
        for layer in self.layers:
            out = layer(out, memory, memory_key_padding_mask, pos, query_pos)
            if self.return_intermediate:
                inter.append(self.norm(out))
        out = self.norm(out)
        if self.return_intermediate:
            inter[-1] = out
            return torch.stack(inter)
        return out.unsqueeze(0)


class OracleTransformer(nn.Module):
    """transformer.py:16-115."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, return_intermediate_dec=False):
        super().__init__()
        self.encoder = OracleTransformerEncoder(d_model, nhead, dim_feedforward, dropout, activation,
                                                normalize_before, num_encoder_layers)
        self.decoder = OracleTransformerDecoder(d_model, nhead, dim_feedforward, dropout, activation,
                                                normalize_before, num_decoder_layers, return_intermediate_dec)
        for p in self.parameters():  # _reset_parameters :57-60
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        self.d_model, self.nhead = d_model, nhead

    def forward(self, src, mask, query_embed, pos_embed, latent_input=None, proprio_input=None,
                additional_pos_embed=None):
        bs = src.shape[0]
        src = src.flatten(2).permute(2, 0, 1)  # (hw, bs, c)
        pos_embed = pos_embed.flatten(2).permute(2, 0, 1)
        if pos_embed.shape[1] == 1:
            pos_embed = pos_embed.repeat(1, bs, 1)
        query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
        additional_pos_embed = additional_pos_embed.unsqueeze(1).repeat(1, bs, 1)
        pos_embed = torch.cat([additional_pos_embed, pos_embed], 0)
        if latent_input.dim() == 2:
            addition_input = torch.stack([latent_input, proprio_input], 0)
        else:
            addition_input = torch.cat([latent_input, proprio_input], 0)
        src = torch.cat([addition_input, src], 0)
        tgt = torch.zeros_like(query_embed)
        memory = self.encoder(src, src_key_padding_mask=mask, pos=pos_embed)
        hs = self.decoder(tgt, memory, memory_key_padding_mask=mask, pos=pos_embed, query_pos=query_embed)
        return hs.transpose(1, 2)


# ----------------------------------------------------------------------------------------------
# ACT policy (point-cloud variants)
# ----------------------------------------------------------------------------------------------
def sinusoid_table(n_position, d_hid):
    """act/utils.py:42-55."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.FloatTensor(ang).unsqueeze(0)


def kl_divergence(mu, logvar):
    """loss/misc.py:11-26 -> total_kld[0]."""
    klds = -0.5 * (1 + logvar - mu.pow(2) - logvar.exp())
    return klds.sum(1).mean(0, True)[0]


def coord_embedding_sine(coord, hidden_dim, temperature=10000):
    """act.py:467-506 (normalize=False path)."""
    npf = hidden_dim // 3
    pad = hidden_dim - npf * 3
    dim_t = torch.arange(npf, dtype=torch.float32, device=coord.device)
    dim_t = temperature ** (2 * (dim_t // 2) / npf)
    parts = []
    for a in range(3):
        p = coord[:, a:a + 1, None] / dim_t
        parts.append(torch.stack((p[..., 0::2].sin(), p[..., 1::2].cos()), dim=2).flatten(1))
    pos = torch.cat(parts, dim=1)
    return torch.cat((pos, torch.zeros_like(pos)[:, :pad]), dim=1)


def oracle_fps(p, o, n_o):
    idx = PO.farthest_point_sampling(p.detach().cpu().numpy(), o.cpu().numpy(), n_o.cpu().numpy())
    return torch.from_numpy(idx).to(p.device)


def oracle_knn(nsample, p, o, n_p, n_o):
    idx, _ = PO.knn_query(nsample, p.detach().cpu().numpy(), o.cpu().numpy(), n_p.detach().cpu().numpy(), n_o.cpu().numpy())
    return torch.from_numpy(idx).to(p.device)


def grouping_with_xyz(idx, feat, xyz, new_xyz):
    """libs/pointops/functions/grouping.py:35-59 (with_xyz=True)."""
    m, ns = idx.shape
    xyz_p = torch.cat([xyz, torch.zeros(1, 3, device=xyz.device)], 0)
    feat_p = torch.cat([feat, torch.zeros(1, feat.shape[1], device=feat.device)], 0)
    flat = idx.reshape(-1).long()
    gf = feat_p[flat].view(m, ns, -1)
    gx = (xyz_p[flat].view(m, ns, 3) - new_xyz.unsqueeze(1)) * torch.sign(idx + 1).unsqueeze(-1)
    return torch.cat((gx, gf), -1)


class OracleACTPCD(nn.Module):
    def __init__(self, backbone, transformer, encoder, hidden_dim, num_queries, num_cameras=0, action_dim=8,
                 qpos_dim=9, env_state_dim=0, latent_dim=32, action_loss=None, klloss=None, kl_weight=20.0,
                 goal_cond_dim=0, obs_feature_pos_embedding=None, freeze_backbone=False, pcd_nsample=16,
                 pcd_npoints=1024, sampling="fps", heatmap_th=0.1, ignore_vae=False, use_mask=False,
                 bg_ratio=0.0, pre_sample=False, in_channels=6):
        super().__init__()
        assert sampling == "fps"  # act.py:443-444 raises for anything else
        self.backbone, self.transformer, self.encoder = backbone, transformer, encoder
        self.hidden_dim, self.num_queries, self.action_dim, self.qpos_dim = hidden_dim, num_queries, action_dim, qpos_dim
        self.latent_dim, self.kl_weight, self.goal_cond_dim = latent_dim, kl_weight, goal_cond_dim
        self.pcd_nsample, self.pcd_npoints, self.ignore_vae = pcd_nsample, pcd_npoints, ignore_vae
        self.pre_sample, self.use_mask, self.bg_ratio = pre_sample, use_mask, bg_ratio
        if freeze_backbone:
            for p in self.backbone.parameters():
                p.requires_grad = False
        # build_encoder (act.py:93-122)
        self.input_proj_robot_state = nn.Linear(qpos_dim, hidden_dim)
        self.cls_embed = nn.Embedding(1, hidden_dim)
        self.encoder_action_proj = nn.Linear(action_dim, hidden_dim)
        self.encoder_joint_proj = nn.Linear(qpos_dim, hidden_dim)
        self.latent_proj = nn.Linear(hidden_dim, latent_dim * 2)
        self.register_buffer("pos_table", sinusoid_table(2 + num_queries, hidden_dim))
        if goal_cond_dim > 0:
            self.proj_goal_cond_emb = nn.Linear(goal_cond_dim, hidden_dim)
        # build_decoder (act.py:124-135)
        self.action_head = nn.Linear(hidden_dim, action_dim)
        self.is_pad_head = nn.Linear(hidden_dim, 1)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.latent_out_proj = nn.Linear(latent_dim, hidden_dim)
        self.additional_pos_embed = nn.Embedding(2 + int(goal_cond_dim > 0), hidden_dim)
        # set abstraction head (act.py:367-379): on the backbone features, or -- pre_sample -- on the raw
        # input channels in front of the backbone
        if not pre_sample:
            self.linear = nn.Linear(3 + backbone.num_channels, hidden_dim, bias=False)
            self.bn = nn.BatchNorm1d(hidden_dim)
        else:
            self.linear = nn.Linear(3 + backbone.in_channels, backbone.in_channels, bias=False)
            self.bn = nn.BatchNorm1d(backbone.in_channels)

    # act.py:137-188
    def forward_encoder(self, d):
        qpos, actions, is_pad = d["qpos"], d.get("actions"), d.get("is_pad")
        bs = qpos.shape[0]
        d["is_training"] = actions is not None
        if d["is_training"] and not self.ignore_vae:
            tok = torch.cat([self.cls_embed.weight.unsqueeze(0).repeat(bs, 1, 1),
                             self.encoder_joint_proj(qpos).unsqueeze(1),
                             self.encoder_action_proj(actions)], 1).permute(1, 0, 2)
            pad = torch.cat([torch.zeros(bs, 2, dtype=torch.bool, device=qpos.device), is_pad], 1)
            pos = self.pos_table.clone().detach().permute(1, 0, 2)
            out = self.encoder(tok, pos=pos, src_key_padding_mask=pad)[0]
            info = self.latent_proj(out)
            mu, logvar = info[:, :self.latent_dim], info[:, self.latent_dim:]
            eps = d.get("_eps")
            if eps is None:
                eps = torch.empty_like(mu).normal_()  # act/utils.py:36-39
            latent = mu + logvar.div(2).exp() * eps
        else:
            mu = logvar = None
            latent = torch.zeros(bs, self.latent_dim, device=qpos.device)
        d["mu"], d["logvar"], d["latent_input"] = mu, logvar, self.latent_out_proj(latent)
        return d

    # act.py:384-465
    def pcd_sampling(self, p, x, o, mask=None):
        b = o.shape[0]
        n_o = torch.arange(1, b + 1, dtype=torch.int32, device=o.device) * self.pcd_npoints
        if not self.use_mask or mask is None:
            idx = oracle_fps(p, o, n_o)
        else:
            # act.py:396-442: FPS separately on the foreground (and, bg_ratio > 0, background) points of
            # the boolean-compacted cloud.  Restated with the reference's quirks: the indices returned
            # by FPS refer to the COMPACTED arrays and are used on the full cloud unchanged, and the
            # background picks are appended after ALL foreground picks (not interleaved per cloud).
            n_bg = int(self.pcd_npoints * self.bg_ratio)
            ar = torch.arange(1, b + 1, dtype=torch.int32, device=o.device)
            fg_n_o = ar * (self.pcd_npoints - n_bg) if self.bg_ratio > 0.0 else n_o
            ends = o.long()
            cm = torch.cumsum(mask.long(), 0)
            fg_o = cm[ends - 1].int()
            fg_idx = oracle_fps(p[mask].contiguous(), fg_o, fg_n_o)
            if self.bg_ratio > 0.0:
                cb = torch.cumsum((~mask).long(), 0)
                bg_idx = oracle_fps(p[~mask].contiguous(), cb[ends - 1].int(), ar * n_bg)
                idx = torch.cat([fg_idx, bg_idx], 0)
            else:
                idx = fg_idx
        n_p = p[idx.long(), :]
        kidx = oracle_knn(self.pcd_nsample, p, o, n_p, n_o)
        g = grouping_with_xyz(kidx, x, p, n_p)  # (m, ns, 3+c)
        y = F.relu(self.bn(self.linear(g).transpose(1, 2).contiguous()))  # (m, c, ns)
        return n_p, y.max(dim=-1).values, n_o, idx

    # act.py:508-598
    def forward_obs_embed(self, d):
        pcd = d["pcds"]
        mask = pcd.get("mask") if self.use_mask else None
        if self.pre_sample:  # act.py:509-527: sample + group the raw channels first, backbone on the M-point cloud
            coord, feats, off, idx = self.pcd_sampling(pcd["coord"], pcd["feat"], pcd["offset"], mask)
            pcd = dict(pcd, coord=coord, feat=feats, offset=off, grid_coord=pcd["grid_coord"][idx.long()])
            feats = self.backbone(pcd)
        else:
            feats = self.backbone(pcd)
            coord, feats, _, _ = self.pcd_sampling(pcd["coord"], feats, pcd["offset"], mask)
        pos = coord_embedding_sine(coord, self.hidden_dim)
        bs = d["qpos"].shape[0]
        src = feats.view(bs, self.pcd_npoints, -1).permute(0, 2, 1).unsqueeze(2)  # (b, c, 1, n)
        pos = pos.view(bs, self.pcd_npoints, -1).permute(0, 2, 1).unsqueeze(2)
        proprio = self.input_proj_robot_state(d["qpos"]).unsqueeze(0)
        if self.goal_cond_dim > 0:
            gc = d["goal_cond"].reshape(bs, -1)
            proprio = torch.cat([proprio, self.proj_goal_cond_emb(gc).unsqueeze(0)], 0)
        d["src"], d["pos"], d["latent_input"], d["proprio_input"] = src, pos, d["latent_input"].unsqueeze(0), proprio
        return d

    def _decode(self, d):
        return self.transformer(d["src"], None, self.query_embed.weight, d["pos"], d["latent_input"],
                                d["proprio_input"], self.additional_pos_embed.weight)[0]

    # act.py:255-279
    def forward_decoder(self, d):
        hs = self._decode(d)
        d["a_hat"], d["is_pad_hat"] = self.action_head(hs), self.is_pad_head(hs)
        return d

    # act.py:281-291
    def forward_loss(self, d):
        kld = kl_divergence(d["mu"], d["logvar"]) if d["mu"] is not None else 0
        al = F.mse_loss(d["a_hat"], d["actions"], reduction="none")
        al = (al * ~d["is_pad"].unsqueeze(-1)).mean()
        d["action_loss"], d["kl_loss"], d["loss"] = al, kld, al + kld * self.kl_weight
        return d

    def forward(self, d):
        d = self.forward_decoder(self.forward_obs_embed(self.forward_encoder(d)))
        return self.forward_loss(d) if d["is_training"] else d


def rotation_6d_to_matrix(d6):
    """src/utils/rotation_conversions.py:556-577 (Zhou et al. 2019): Gram-Schmidt on the two 3-vectors."""
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def matrix_to_quaternion(m):
    """src/utils/rotation_conversions.py:102-161 restated element by element: the four |q_i| from the diagonal,
    the candidate built around the largest one (best conditioned), then standardised to a non-negative real part."""
    flat = m.reshape(-1, 3, 3)
    out = torch.empty(flat.shape[0], 4, dtype=m.dtype)
    for n, r in enumerate(flat):
        (m00, m01, m02), (m10, m11, m12), (m20, m21, m22) = [[float(v) for v in row] for row in r]
        q = [max(0.0, 1 + m00 + m11 + m22) ** 0.5, max(0.0, 1 + m00 - m11 - m22) ** 0.5,
             max(0.0, 1 - m00 + m11 - m22) ** 0.5, max(0.0, 1 - m00 - m11 + m22) ** 0.5]
        rows = [[q[0] ** 2, m21 - m12, m02 - m20, m10 - m01], [m21 - m12, q[1] ** 2, m10 + m01, m02 + m20],
                [m02 - m20, m10 + m01, q[2] ** 2, m12 + m21], [m10 - m01, m20 + m02, m21 + m12, q[3] ** 2]]
        i = max(range(4), key=lambda j: (q[j], -j))  # argmax, first index on ties (torch.argmax)
        cand = [v / (2.0 * max(q[i], 0.1)) for v in rows[i]]
        if cand[0] < 0:
            cand = [-v for v in cand]
        out[n] = torch.tensor(cand, dtype=m.dtype)
    return out.reshape(m.shape[:-2] + (4,))


class OracleACTRLBenchPCD(OracleACTPCD):
    """act.py:707-825 (training: rot6d kept raw, sigmoid gripper / collision; inference: rot6d -> quaternion)."""

    def __init__(self, *args, rot_type="6d", collision=False, position_loss_weight=1.0, **kwargs):
        super().__init__(*args, **kwargs)
        self.rot_type, self.collision, self.position_loss_weight = rot_type, collision, position_loss_weight

    def forward_decoder(self, d):
        hs = self._decode(d)
        a = self.action_head(hs)
        position = a[..., :3]
        if self.collision:
            gripper = torch.cat([torch.sigmoid(a[..., -2:-1]), torch.sigmoid(a[..., -1:])], -1)
            rot = a[..., 3:-2]
        else:
            gripper = torch.sigmoid(a[..., -1:])
            rot = a[..., 3:-1]
        if not d["is_training"]:  # act.py:785-795: rot6d -> matrix -> quaternion (w first, w >= 0)
            assert self.rot_type == "6d"
            rot = matrix_to_quaternion(rotation_6d_to_matrix(rot))
        d["a_hat"], d["is_pad_hat"] = torch.cat([position, rot, gripper], -1), self.is_pad_head(hs)
        return d

    def forward_loss(self, d):
        kld = kl_divergence(d["mu"], d["logvar"])
        al = F.mse_loss(d["a_hat"], d["actions"], reduction="none")
        al = torch.cat([al[..., :3] * self.position_loss_weight, al[..., 3:]], -1)
        al = (al * ~d["is_pad"].unsqueeze(-1)).mean()
        d["action_loss"], d["kl_loss"], d["loss"] = al, kld, al + kld * self.kl_weight
        return d


def build_oracle_policy(cfg: dict, rlbench: bool = False):
    """cfg keys: hidden_dim, nhead, dim_feedforward, enc_layers, dec_layers, dropout, num_queries,
    action_dim, qpos_dim, goal_cond_dim, latent_dim, kl_weight, pcd_npoints, pcd_nsample, in_channels."""
    backbone = OraclePointNet(cfg.get("in_channels", 6), int(cfg.get("backbone_classes", 0)))
    tr = OracleTransformer(cfg["hidden_dim"], cfg["nhead"], cfg["enc_layers"], cfg["dec_layers"],
                           cfg["dim_feedforward"], cfg["dropout"], "relu", False, True)
    enc = OracleTransformerEncoder(cfg["hidden_dim"], cfg["nhead"], cfg["dim_feedforward"], cfg["dropout"], "relu",
                                   False, cfg["enc_layers"])
    cls = OracleACTRLBenchPCD if rlbench else OracleACTPCD
    extra = dict(collision=cfg.get("collision", False), position_loss_weight=cfg.get("position_loss_weight", 1.0)) if rlbench else {}
    return cls(backbone, tr, enc, cfg["hidden_dim"], cfg["num_queries"], 0, cfg["action_dim"], cfg["qpos_dim"],
               latent_dim=cfg.get("latent_dim", 32), kl_weight=cfg.get("kl_weight", 10.0),
               goal_cond_dim=cfg.get("goal_cond_dim", 0), pcd_nsample=cfg.get("pcd_nsample", 16),
               pcd_npoints=cfg["pcd_npoints"], use_mask=bool(cfg.get("use_mask", False)),
               bg_ratio=float(cfg.get("bg_ratio", 0.0)), pre_sample=bool(cfg.get("pre_sample", False)), **extra)
