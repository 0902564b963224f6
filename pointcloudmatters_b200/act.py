"""ACT policy for point-cloud observations -- host-side mirror of
`src/models/components/act/act.py` (`ACTPCD` :312-598, `ACTRLBenchPCD` :707-825, base `ACT`
:40-309) of HaoyiZhu/PointCloudMatters.  Same constructor kwargs, same `state_dict` keys and
shapes, same `forward(data_dict) -> data_dict` contract (keys `loss, action_loss, kl_loss, a_hat,
is_pad_hat, mu, logvar`), so it can be selected with a one-line Hydra `_target_` override.

What changed underneath (B200-first):
  * FPS / kNN run through the sm_100a kernels of libpcm_b200.so; the per-cloud Python loops and
    host syncs of the reference (`functions/sampling.py:14-17`, act.py:387-391) are gone: the
    fixed-M `new_offset` is built on the device and `pcds["n_max"]` (optional Python int, the
    largest cloud) makes the whole step sync-free;
  * the set-abstraction head (gather -> Linear -> BatchNorm -> ReLU -> max) is one fused operator;
  * dense blocks go through pointcloudmatters_b200.functional.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as PF
from . import pointops


def get_sinusoid_encoding_table(n_position, d_hid):
    """act/utils.py:42-55."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table).unsqueeze(0)


class KLDivergence(nn.Module):
    """loss/misc.py:6-26."""

    def forward(self, mu, logvar):
        if mu is None:
            return 0
        klds = -0.5 * (1 + logvar - mu.pow(2) - logvar.exp())
        return klds.sum(1).mean(0, True)[0]


def sample_indices(p, o32, npoints, mask, bg_ratio, hints):
    """FPS picks of the set-abstraction front end, shared by `ACTPCD.pcd_sampling` (act.py:393-442) and
    `PCDObsEncoder.pcd_sampling` (pcd_obs_encoder.py:123-177).  Plain (mask is None): one FPS over the batch.
    Masked: FPS separately on the foreground (and, bg_ratio > 0, background) points of the boolean-compacted
    clouds, with the reference's quirks kept as they are: the indices FPS returns refer to the COMPACTED arrays
    and are applied to the full cloud unchanged, and all background picks follow all foreground picks
    (`torch.cat([fg_idx, bg_idx])`, not re-interleaved per cloud).

    The compaction is a stable argsort instead of `p[mask]` (static shapes, no host sync; FPS reads only the rows
    below its offsets) and the per-cloud counts are one cumsum instead of the reference's O(B) `.sum().item()`
    loops (act.py:423-426,434-437).  `hints`: host-known largest cloud sizes `n_max` / `fg_n_max` / `bg_n_max`
    (Python ints) make the call free of device->host reads."""
    b = o32.shape[0]
    if mask is None:
        return pointops.farthest_point_sampling(p, o32, torch.arange(1, b + 1, dtype=torch.int32, device=p.device) * npoints,
                                                n_max=hints.get("n_max"), m_total=b * npoints)
    n_bg = int(npoints * bg_ratio) if bg_ratio > 0.0 else 0
    ar = torch.arange(1, b + 1, dtype=torch.int32, device=p.device)
    ends = o32.long() - 1
    picks = []
    for keep, per, hint in ((mask, npoints - n_bg, "fg_n_max"), (~mask, n_bg, "bg_n_max")):
        if per == 0:
            continue
        order = torch.argsort(~keep, stable=True)  # kept rows first, original order preserved
        sub_o = torch.cumsum(keep.int(), 0, dtype=torch.int32)[ends].contiguous()
        picks.append(pointops.farthest_point_sampling(p[order].contiguous(), sub_o, ar * per,
                                                      n_max=hints.get(hint), m_total=b * per))
    return picks[0] if len(picks) == 1 else torch.cat(picks, 0)


class ACTPCD(nn.Module):
    def __init__(self, backbone, transformer, encoder, hidden_dim, num_queries, num_cameras=0, action_dim=8,
                 qpos_dim=9, env_state_dim=0, latent_dim=32, action_loss=None, klloss=None, kl_weight=20.0,
                 goal_cond_dim=0, obs_feature_pos_embedding=None, freeze_backbone=False, pcd_nsample=16,
                 pcd_npoints=1024, sampling="fps", heatmap_th=0.1, ignore_vae=False, use_mask=False,
                 bg_ratio=0.0, pre_sample=False, in_channels=6):
        super().__init__()
        if backbone is None:
            raise NotImplementedError("state-only ACT (backbone=None) is outside the point-cloud hot path")
        if "fps" not in sampling:
            raise NotImplementedError(sampling)  # same as the reference (act.py:443-444)
        self.backbone, self.transformer, self.encoder = backbone, transformer, encoder
        self.num_queries, self.num_cameras = num_queries, 0
        self.action_dim, self.qpos_dim, self.env_state_dim = action_dim, qpos_dim, env_state_dim
        self.hidden_dim, self.kl_weight, self.latent_dim = hidden_dim, kl_weight, latent_dim
        self.goal_cond_dim, self.freeze_backbone, self.ignore_vae = goal_cond_dim, freeze_backbone, ignore_vae
        self.obs_feature_pos_embedding = None
        if freeze_backbone:
            for p in self.backbone.parameters():
                p.requires_grad = False
        self.action_loss = action_loss if action_loss is not None else nn.MSELoss(reduction="none")
        self.klloss = klloss if klloss is not None else KLDivergence()
        # build_encoder (act.py:93-122)
        self.input_proj_robot_state = nn.Linear(qpos_dim, hidden_dim)
        self.cls_embed = nn.Embedding(1, hidden_dim)
        self.encoder_action_proj = nn.Linear(action_dim, hidden_dim)
        self.encoder_joint_proj = nn.Linear(qpos_dim, hidden_dim)
        self.latent_proj = nn.Linear(hidden_dim, latent_dim * 2)
        self.register_buffer("pos_table", get_sinusoid_encoding_table(2 + num_queries, hidden_dim))
        if goal_cond_dim > 0:
            self.proj_goal_cond_emb = nn.Linear(goal_cond_dim, hidden_dim)
        # build_decoder (act.py:124-135)
        self.action_head = nn.Linear(hidden_dim, action_dim)
        self.is_pad_head = nn.Linear(hidden_dim, 1)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.latent_out_proj = nn.Linear(latent_dim, hidden_dim)
        self.additional_pos_embed = nn.Embedding(2 + int(goal_cond_dim > 0), hidden_dim)
        self.input_proj = None
        # set-abstraction head (act.py:363-382)
        self.pcd_nsample, self.pcd_npoints, self.pre_sample = pcd_nsample, pcd_npoints, pre_sample
        if not pre_sample:
            self.linear = nn.Linear(3 + self.backbone.num_channels, hidden_dim, bias=False)
            self.bn = nn.BatchNorm1d(hidden_dim)
        else:  # act.py:371-375: the head runs on the raw input channels, in front of the backbone
            self.linear = nn.Linear(3 + self.backbone.in_channels, self.backbone.in_channels, bias=False)
            self.bn = nn.BatchNorm1d(self.backbone.in_channels)
        self.pool = nn.MaxPool1d(pcd_nsample)
        self.relu = nn.ReLU(inplace=True)
        self.sampling, self.use_mask, self.bg_ratio = sampling, use_mask, bg_ratio

    # ---- CVAE posterior (act.py:137-188) ------------------------------------------------------
    def forward_encoder(self, data_dict):
        qpos = data_dict["qpos"]
        actions = data_dict.get("actions", None)
        is_pad = data_dict.get("is_pad", None)
        is_training = actions is not None
        bs = qpos.shape[0]
        data_dict["is_training"] = is_training
        if is_training and not self.ignore_vae:
            action_embed = PF.linear(actions, self.encoder_action_proj.weight, self.encoder_action_proj.bias)
            qpos_embed = PF.linear(qpos, self.encoder_joint_proj.weight, self.encoder_joint_proj.bias).unsqueeze(1)
            cls_embed = self.cls_embed.weight.unsqueeze(0).expand(bs, -1, -1)
            encoder_input = torch.cat([cls_embed, qpos_embed, action_embed], dim=1).permute(1, 0, 2).contiguous()
            pad = torch.cat([torch.zeros(bs, 2, dtype=torch.bool, device=qpos.device), is_pad], dim=1)
            pos_embed = self.pos_table.detach().permute(1, 0, 2)
            cls_out = self.encoder(encoder_input, pos=pos_embed, src_key_padding_mask=pad)[0]
            latent_info = PF.linear(cls_out, self.latent_proj.weight, self.latent_proj.bias)
            mu, logvar = latent_info[:, : self.latent_dim], latent_info[:, self.latent_dim:]
            eps = data_dict.get("_eps", None)  # test hook: injected reparametrisation noise
            if eps is None:
                eps = torch.empty_like(mu).normal_()
            latent_sample = mu + logvar.div(2).exp() * eps
        else:
            mu = logvar = None
            latent_sample = torch.zeros(bs, self.latent_dim, dtype=torch.float32, device=qpos.device)
        data_dict["mu"], data_dict["logvar"] = mu, logvar
        data_dict["latent_input"] = PF.linear(latent_sample, self.latent_out_proj.weight, self.latent_out_proj.bias)
        return data_dict

    # ---- set abstraction (act.py:384-465) ----------------------------------------------------
    def _sample_indices(self, p, o32, n_o, mask, hints):
        return sample_indices(p, o32, self.pcd_npoints, mask if self.use_mask else None, self.bg_ratio, hints)

    def _n_head_rows(self):
        """Non-point rows in front of the encoder sequence: latent, proprio (+ goal) (transformer.py:89-92)."""
        return 2 + int(self.goal_cond_dim > 0)

    def _token_fast_path(self, feat):
        """True when the set-abstraction head can write the transformer's (S, B, E) token tensor directly (and the sine
        embedding its positional tensor): our own Transformer, fused-kernel widths, CUDA, head in front of the encoder."""
        return (not self.pre_sample and feat.is_cuda and getattr(self.transformer, "accepts_token_buffers", False)
                and self.hidden_dim % 4 == 0 and self.linear.weight.shape[0] == self.hidden_dim
                and feat.shape[1] % 8 == 0 and self.hidden_dim % 8 == 0)

    def coord_embedding_sine_tokens(self, coord, b, temperature=10000):
        """`coord_embedding_sine` of the sampled coordinates, written by ONE kernel straight into the transformer's
        positional tensor (S, B, E) = [additional_pos_embed rows (broadcast over the batch) ; sine rows] -- replaces the
        dozen elementwise kernels below plus the flatten / permute / repeat / cat passes of transformer.py:75-88."""
        from ._lib import check, current_stream, lib, ptr

        E, head = self.hidden_dim, self._n_head_rows()
        npf = E // 3
        dt = getattr(self, "_dim_t", None)
        if dt is None or dt.device != coord.device or dt.numel() != npf:
            j = torch.arange(npf, dtype=torch.float32, device=coord.device)
            dt = self._dim_t = (temperature ** (2 * (j // 2) / npf)).contiguous()
        per = coord.shape[0] // b
        pos = torch.empty((head + per, b, E), dtype=torch.float32, device=coord.device)
        add = self.additional_pos_embed.weight.detach()
        check(lib.pcm_coord_embed_sine_tokens(per, b, head, E, npf, ptr(coord.contiguous()), ptr(dt), ptr(add), ptr(pos),
                                              current_stream()), "pcm_coord_embed_sine_tokens")
        pos._pcm_add_pos = self.additional_pos_embed.weight
        return pos

    def pcd_sampling(self, pxo, mask=None, return_index=False, n_max=None, hints=None, tokens=False):
        p, x, o = pxo
        b = o.shape[0]
        pre = getattr(self, "_presampled", None)
        if pre is not None:
            side, n_o, o32, idx, n_p, knn_idx = pre
            torch.cuda.current_stream().wait_stream(side)  # join
            for t in (n_o, idx, n_p, knn_idx, getattr(self, "_presampled_pos", None)):
                if t is not None:
                    t.record_stream(torch.cuda.current_stream())
            self._presampled = None
        else:
            hints = dict(hints or {})
            hints.setdefault("n_max", n_max)
            n_o = torch.arange(1, b + 1, dtype=torch.int32, device=o.device) * self.pcd_npoints
            o32 = o.int() if o.dtype != torch.int32 else o
            idx = self._sample_indices(p, o32, n_o, mask, hints)
            n_p = p[idx.long(), :].contiguous()
            knn_idx, _ = pointops.ops.KNNQuery.apply(self.pcd_nsample, p, o32, n_p, n_o, False)
        # cloud-size bound for the set-abstraction kernels: only the plain (unmasked) sampling path has one for ALL points
        sa_n_max = (hints or {}).get("n_max", None) or n_max
        if mask is not None or not isinstance(sa_n_max, int):
            sa_n_max = None
        if tokens:
            pos = getattr(self, "_presampled_pos", None)
            if pos is None or pos.dim() != 3:
                pos = self._presampled_pos = self.coord_embedding_sine_tokens(n_p, b)
            x = PF.set_abstraction(p, x, o32, n_p, n_o, knn_idx, self.linear.weight, self.bn,
                                   tokens=(b, self._n_head_rows(), pos), n_max=sa_n_max)
        else:
            x = PF.set_abstraction(p, x, o32, n_p, n_o, knn_idx, self.linear.weight, self.bn, n_max=sa_n_max)
        if return_index:
            return [n_p, x, n_o, idx]
        return [n_p, x, n_o]

    def coord_embedding_sine(self, coord, temperature=10000):
        """act.py:467-506 (normalize=False, the only form the reference calls)."""
        npf = self.hidden_dim // 3
        pad = self.hidden_dim - npf * 3
        dim_t = torch.arange(npf, dtype=torch.float32, device=coord.device)
        dim_t = temperature ** (2 * (dim_t // 2) / npf)
        parts = []
        for a in range(3):
            pa = coord[:, a:a + 1, None] / dim_t
            parts.append(torch.stack((pa[..., 0::2].sin(), pa[..., 1::2].cos()), dim=2).flatten(1))
        pos = torch.cat(parts, dim=1)
        return torch.cat((pos, torch.zeros_like(pos)[:, :pad]), dim=1)

    def forward_pcd_embed(self, pcd_dict):
        mask = pcd_dict.get("mask", None) if self.use_mask else None
        hints = {k: pcd_dict.get(k, None) for k in ("n_max", "fg_n_max", "bg_n_max")}
        if self.pre_sample:
            # act.py:509-527: sample + group the RAW channels first (Linear(3+c -> c)), re-index the voxel
            # coordinates, then run the backbone on the M-point cloud
            coord, features, offset, idx = self.pcd_sampling((pcd_dict["coord"], pcd_dict["feat"], pcd_dict["offset"]),
                                                             mask, return_index=True, hints=hints)
            pcd_dict["coord"], pcd_dict["feat"], pcd_dict["offset"] = coord, features, offset
            pcd_dict["grid_coord"] = pcd_dict["grid_coord"][idx.long()]
            features = self.backbone(pcd_dict)
        else:
            features = self.backbone(pcd_dict)
            if self._token_fast_path(features):
                # the head writes the (S, B, E) token tensor itself; hand the transformer (b, c, 1, n) VIEWS of the point
                # rows (the reference's shapes) that carry the full buffers (`_pcm_tokens`)
                coord, tok, _ = self.pcd_sampling((pcd_dict["coord"], features, pcd_dict["offset"]), mask, hints=hints, tokens=True)
                pos_tok = self._presampled_pos
                self._presampled_pos = None
                head = self._n_head_rows()
                fv = tok[head:].permute(1, 2, 0).unsqueeze(2)
                pv = pos_tok[head:].permute(1, 2, 0).unsqueeze(2)
                fv._pcm_tokens, pv._pcm_tokens = tok, pos_tok
                return fv, pv
            coord, features, _ = self.pcd_sampling((pcd_dict["coord"], features, pcd_dict["offset"]), mask, hints=hints)
        pcd_pos = getattr(self, "_presampled_pos", None)  # computed on the FPS / kNN side stream when forked
        self._presampled_pos = None
        if pcd_pos is None or pcd_pos.dim() != 2:
            pcd_pos = self.coord_embedding_sine(coord)
        b = pcd_dict["offset"].shape[0]
        features = features.view(b, self.pcd_npoints, -1).permute(0, 2, 1).unsqueeze(2)  # (b, c, 1, n)
        pcd_pos = pcd_pos.view(b, self.pcd_npoints, -1).permute(0, 2, 1).unsqueeze(2)
        return features, pcd_pos

    # ---- observation tokens (act.py:553-598) -------------------------------------------------
    def forward_obs_embed(self, data_dict):
        qpos = data_dict["qpos"]
        latent_input = data_dict["latent_input"]
        pcd_tokens, pcd_pos = self.forward_pcd_embed(data_dict["pcds"])
        proprio_input = PF.linear(qpos, self.input_proj_robot_state.weight, self.input_proj_robot_state.bias).unsqueeze(0)
        if self.goal_cond_dim > 0:
            gc = data_dict["goal_cond"]
            if gc.dim() > 2:
                gc = gc.reshape(gc.shape[0], -1)
                data_dict["goal_cond"] = gc
            goal = PF.linear(gc, self.proj_goal_cond_emb.weight, self.proj_goal_cond_emb.bias).unsqueeze(0)
            proprio_input = torch.cat([proprio_input, goal], dim=0)
        data_dict["src"], data_dict["pos"] = pcd_tokens, pcd_pos
        data_dict["latent_input"], data_dict["proprio_input"] = latent_input.unsqueeze(0), proprio_input
        return data_dict

    def _decode(self, data_dict):
        args = (data_dict["src"], None, self.query_embed.weight, data_dict["pos"], data_dict["latent_input"],
                data_dict["proprio_input"], self.additional_pos_embed.weight)
        if getattr(self.transformer, "accepts_token_buffers", False):  # our Transformer: hs[0] without building the stack
            return self.transformer(*args, first_only=True)
        return self.transformer(*args)[0]

    # ---- heads + loss (act.py:255-291) -------------------------------------------------------
    def forward_decoder(self, data_dict):
        hs = data_dict.pop("_hs", None)  # already decoded by a `_fused_heads` attempt that fell back
        if hs is None:
            hs = self._decode(data_dict)
        data_dict["a_hat"] = PF.linear(hs, self.action_head.weight, self.action_head.bias)
        data_dict["is_pad_hat"] = PF.linear(hs, self.is_pad_head.weight, self.is_pad_head.bias)
        return data_dict

    def forward_loss(self, data_dict):
        total_kld = self.klloss(data_dict["mu"], data_dict["logvar"])
        action_loss = self.action_loss(data_dict["a_hat"], data_dict["actions"])
        action_loss = (action_loss * ~data_dict["is_pad"].unsqueeze(-1)).mean()
        data_dict["action_loss"], data_dict["kl_loss"] = action_loss, total_kld
        data_dict["loss"] = action_loss + total_kld * self.kl_weight
        return data_dict

    def _head_cfg(self):
        """(sigmoid start index, position dims, position loss weight) of the action head; ManiSkill: plain MSE."""
        return self.action_dim, 0, 1.0

    def _fused_heads(self, data_dict):
        """Heads (+ loss when training) as ONE kernel each way (csrc/tokens.cu); returns False when the configuration
        needs the composed path (custom loss modules, widths the kernel does not cover)."""
        if not (type(self.action_loss) is nn.MSELoss and self.action_loss.reduction == "none"
                and type(self.klloss).__name__ == "KLDivergence"):
            return False
        hs = self._decode(data_dict)
        data_dict["_hs"] = hs
        if not hs.is_cuda:
            return False
        training = data_dict["is_training"]
        sig, n_pos, w_pos = self._head_cfg()
        out = PF.act_heads_loss(hs, self.action_head, self.is_pad_head, data_dict["actions"] if training else None,
                                data_dict["is_pad"] if training else None, data_dict["mu"] if training else None,
                                data_dict["logvar"] if training else None, self.kl_weight, sig, n_pos, w_pos)
        if out is None:
            return False
        data_dict["a_hat"], data_dict["is_pad_hat"] = out[0], out[1]
        if training:
            data_dict["loss"], data_dict["action_loss"], data_dict["kl_loss"] = out[2], out[3], out[4]
        data_dict.pop("_hs", None)
        return True

    def grad_buckets(self):
        """Parameter-name prefixes of the gradient buckets in the order their gradients become final during backward, with
        the boundary tag (functional.grad_boundary) that closes each; the last bucket (None) is everything else and closes
        when backward returns.  The trainer lays the flat gradient out in this order and all-reduces bucket by bucket."""
        return [("transformer.decoder", ("transformer.decoder.", "action_head.", "is_pad_head.")),
                ("transformer.encoder", ("transformer.encoder.",)),
                (None, ("",))]

    def _presample(self, data_dict):
        """FPS + kNN depend only on the input coordinates and are latency-bound (a chain of M-1
        dependent rounds on 64 of the 148 SMs): run them on a side stream while the CVAE encoder
        occupies the main stream.  Under CUDA-graph capture this becomes a fork/join in the graph."""
        pcd = data_dict["pcds"]
        p, o = pcd["coord"], pcd["offset"]
        if not p.is_cuda:
            return
        b = o.shape[0]
        main = torch.cuda.current_stream()
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream()
        side = self._side_stream
        side.wait_stream(main)
        with torch.cuda.stream(side), PF.stage("fps+knn+sine (side stream)"):
            n_o = torch.arange(1, b + 1, dtype=torch.int32, device=o.device) * self.pcd_npoints
            o32 = o.int() if o.dtype != torch.int32 else o
            hints = {k: pcd.get(k, None) for k in ("n_max", "fg_n_max", "bg_n_max")}
            idx = self._sample_indices(p, o32, n_o, pcd.get("mask", None) if self.use_mask else None, hints)
            n_p = p[idx.long(), :].contiguous()
            knn_idx, _ = pointops.ops.KNNQuery.apply(self.pcd_nsample, p, o32, n_p, n_o, False)
            # the sine embedding of the sampled coordinates depends on n_p only
            if (not self.pre_sample and getattr(self.transformer, "accepts_token_buffers", False)
                    and self.linear.weight.shape[0] == self.hidden_dim and self.hidden_dim % 8 == 0):
                self._presampled_pos = self.coord_embedding_sine_tokens(n_p, b)
            else:
                self._presampled_pos = self.coord_embedding_sine(n_p)
        self._presampled = (side, n_o, o32, idx, n_p, knn_idx)

    def sync_free(self, pcds) -> bool:
        """True when the host-known cloud-size hints make FPS run without a device->host read."""
        need = ["n_max"] if not (self.use_mask and pcds.get("mask", None) is not None) else (
            ["fg_n_max"] + (["bg_n_max"] if self.bg_ratio > 0.0 else []))
        return all(pcds.get(k, None) is not None for k in need)

    def _encode_on_side_stream(self, data_dict):
        """CVAE posterior (`forward_encoder`) on its own stream.  It depends only on qpos / actions / is_pad and
        is a chain of ~50 small kernels (102 tokens x B rows: a third of the SMs at best), independent of the
        PointNet -> FPS/kNN -> set-abstraction chain that produces the observation tokens: the two run side by
        side until the transformer needs `latent_input`.  Autograd replays each node on the stream of its
        forward, so the backward passes of the two chains overlap as well; under CUDA-graph capture both become
        parallel branches of the step graph."""
        main = torch.cuda.current_stream()
        if getattr(self, "_enc_stream", None) is None:
            self._enc_stream = torch.cuda.Stream()
        side = self._enc_stream
        side.wait_stream(main)
        with torch.cuda.stream(side), PF.stage("cvae encoder (side stream)"):
            data_dict = self.forward_encoder(data_dict)
        return data_dict, side

    def forward(self, data_dict):
        self._presampled = self._presampled_pos = None
        fork = self.sync_free(data_dict["pcds"]) and data_dict["qpos"].is_cuda
        if fork:
            self._presample(data_dict)
            data_dict, enc_stream = self._encode_on_side_stream(data_dict)
        else:
            with PF.stage("cvae encoder"):
                data_dict = self.forward_encoder(data_dict)
        with PF.stage("pointnet + set abstraction"):
            data_dict = self.forward_obs_embed(data_dict)
        if fork:  # join: the transformer consumes latent_input, the loss mu / logvar
            main = torch.cuda.current_stream()
            main.wait_stream(enc_stream)
            for k in ("latent_input", "mu", "logvar"):
                if torch.is_tensor(data_dict.get(k, None)):
                    data_dict[k].record_stream(main)
        with PF.stage("transformer + heads + loss"):
            fused = self._fused_heads(data_dict)
        if fused:
            return self._finish_inference(data_dict) if not data_dict["is_training"] else data_dict
        data_dict = self.forward_decoder(data_dict)
        if not data_dict["is_training"]:
            return data_dict
        return self.forward_loss(data_dict)

    def _finish_inference(self, data_dict):
        return data_dict


class ACTRLBenchPCD(ACTPCD):
    """act.py:707-825: position / rot6d / sigmoid(gripper[, collision]) heads, weighted position loss."""

    def __init__(self, *args, rot_type="6d", collision=False, position_loss_weight=1.0, **kwargs):
        super().__init__(*args, **kwargs)
        self.rot_type, self.collision, self.position_loss_weight = rot_type, collision, position_loss_weight

    def _head_cfg(self):
        return self.action_dim - (2 if self.collision else 1), 3, float(self.position_loss_weight)

    def _finish_inference(self, data_dict):
        """act.py:785-795 after the fused heads (sigmoids already applied): rot6d -> quaternion."""
        if self.rot_type != "6d":
            raise NotImplementedError
        a = data_dict["a_hat"]
        sig = self._head_cfg()[0]
        rot = _matrix_to_quaternion(_rotation_6d_to_matrix(a[..., 3:sig]))
        data_dict["a_hat"] = torch.cat([a[..., :3], rot, a[..., sig:]], dim=-1)
        return data_dict

    def forward_decoder(self, data_dict):
        hs = data_dict.pop("_hs", None)
        if hs is None:
            hs = self._decode(data_dict)
        a_hat = PF.linear(hs, self.action_head.weight, self.action_head.bias)
        position = a_hat[..., :3]
        if self.collision:
            gripper = torch.cat([torch.sigmoid(a_hat[..., -2:-1]), torch.sigmoid(a_hat[..., -1:])], dim=-1)
            rot = a_hat[..., 3:-2]
        else:
            gripper = torch.sigmoid(a_hat[..., -1:])
            rot = a_hat[..., 3:-1]
        if not data_dict["is_training"]:
            if self.rot_type != "6d":
                raise NotImplementedError
            rot = _matrix_to_quaternion(_rotation_6d_to_matrix(rot))
        data_dict["a_hat"] = torch.cat([position, rot, gripper], dim=-1)
        data_dict["is_pad_hat"] = PF.linear(hs, self.is_pad_head.weight, self.is_pad_head.bias)
        return data_dict

    def forward_loss(self, data_dict):
        total_kld = self.klloss(data_dict["mu"], data_dict["logvar"])
        action_loss = self.action_loss(data_dict["a_hat"], data_dict["actions"])
        action_loss = torch.cat([action_loss[..., :3] * self.position_loss_weight, action_loss[..., 3:]], dim=-1)
        action_loss = (action_loss * ~data_dict["is_pad"].unsqueeze(-1)).mean()
        data_dict["action_loss"], data_dict["kl_loss"] = action_loss, total_kld
        data_dict["loss"] = action_loss + total_kld * self.kl_weight
        return data_dict


# rot6d -> matrix -> quaternion for the (non-training) RLBench inference branch
# (reference: src/utils/rotation_conversions.py, Zhou et al. 2019 / standard w-first quaternion).
def _rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1, dim=-1)
    return torch.stack((b1, b2, torch.cross(b1, b2, dim=-1)), dim=-2)


def _matrix_to_quaternion(matrix):
    m = matrix
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    q_abs = torch.sqrt(torch.clamp(torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22,
                                                1 - m00 + m11 - m22, 1 - m00 - m11 + m22], dim=-1), min=0))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1]], -1),
        torch.stack([m[..., 2, 1] - m[..., 1, 2], q_abs[..., 1] ** 2, m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0]], -1),
        torch.stack([m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] + m[..., 0, 1], q_abs[..., 2] ** 2, m[..., 1, 2] + m[..., 2, 1]], -1),
        torch.stack([m[..., 1, 0] - m[..., 0, 1], m[..., 2, 0] + m[..., 0, 2], m[..., 2, 1] + m[..., 1, 2], q_abs[..., 3] ** 2], -1),
    ], dim=-2)
    cand = cand / (2.0 * q_abs[..., None].clamp(min=0.1))
    best = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    out = cand[best, :].reshape(matrix.shape[:-2] + (4,))
    return torch.where(out[..., 0:1] < 0, -out, out)  # standardize_quaternion (rotation_conversions.py:368-380): w >= 0


def build_policy(cfg: dict, rlbench: bool = False):
    """Convenience constructor with the reference's maniskill2_act_pcd_model.yaml structure."""
    from .pointnet import PointNet
    from .transformer import Transformer, TransformerEncoder

    backbone = PointNet(cfg.get("in_channels", 6), int(cfg.get("backbone_classes", 0)))
    tr = Transformer(d_model=cfg["hidden_dim"], nhead=cfg["nhead"], num_encoder_layers=cfg["enc_layers"],
                     num_decoder_layers=cfg["dec_layers"], dim_feedforward=cfg["dim_feedforward"],
                     dropout=cfg["dropout"], normalize_before=False, return_intermediate_dec=True)
    enc = TransformerEncoder(d_model=cfg["hidden_dim"], nhead=cfg["nhead"], dim_feedforward=cfg["dim_feedforward"],
                             dropout=cfg["dropout"], num_layers=cfg["enc_layers"], normalize_before=False)
    kw = dict(backbone=backbone, transformer=tr, encoder=enc, hidden_dim=cfg["hidden_dim"],
              num_queries=cfg["num_queries"], num_cameras=1, action_dim=cfg["action_dim"], qpos_dim=cfg["qpos_dim"],
              latent_dim=cfg.get("latent_dim", 32), kl_weight=cfg.get("kl_weight", 10.0),
              goal_cond_dim=cfg.get("goal_cond_dim", 0), pcd_nsample=cfg.get("pcd_nsample", 16),
              pcd_npoints=cfg["pcd_npoints"], use_mask=bool(cfg.get("use_mask", False)),
              bg_ratio=float(cfg.get("bg_ratio", 0.0)), pre_sample=bool(cfg.get("pre_sample", False)))
    if rlbench:
        return ACTRLBenchPCD(**kw, collision=cfg.get("collision", False),
                             position_loss_weight=cfg.get("position_loss_weight", 1.0))
    return ACTPCD(**kw)


# BASELINE.json configs: model hyper-parameters of maniskill2_act_pcd_model.yaml
ACT_MODEL_CFG = dict(hidden_dim=512, nhead=8, dim_feedforward=32, enc_layers=4, dec_layers=7, dropout=0.1,
                     num_queries=100, latent_dim=32, kl_weight=10.0, pcd_nsample=16)
