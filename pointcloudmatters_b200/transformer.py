"""DETR-style transformer of ACT -- host-side mirror of
`src/models/components/act/transformer.py:16-425` (same class names, constructor kwargs, parameter
names; `nn.MultiheadAttention` / `nn.Linear` / `nn.LayerNorm` are kept purely as PARAMETER
CONTAINERS so reference checkpoints load key-for-key -- their forward is never called; all
arithmetic goes through pointcloudmatters_b200.functional -> libpcm_b200.so).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import functional as PF


def _check_activation(activation):
    if activation != "relu":
        raise NotImplementedError("the B200 path implements the reference's configured activation (relu) only")


class TransformerEncoderLayer(nn.Module):
    """transformer.py:210-283."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        _check_activation(activation)
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.normalize_before = normalize_before
        self.p = dropout

    def _ffn(self, x):
        return PF.feed_forward(x, self.linear1, self.linear2, self.p, self.training)

    def forward(self, src, src_mask: Optional[Tensor] = None, src_key_padding_mask: Optional[Tensor] = None,
                pos: Optional[Tensor] = None, pos_head: Optional[Tensor] = None):
        """`pos_head` (extension, see functional.multi_head_attention): learned leading rows of a
        detached `pos`."""
        assert src_mask is None, "attn_mask is never used by the reference ACT path"
        tr = self.training
        if self.normalize_before:
            s2 = PF.add_dropout_layernorm(None, src, self.norm1, 0.0, False)
            src = src + PF.dropout(PF.multi_head_attention(self.self_attn, s2, pos, None, None, src_key_padding_mask, tr,
                                                           pos_head=pos_head), self.p, tr)
            s2 = PF.add_dropout_layernorm(None, src, self.norm2, 0.0, False)
            return src + PF.dropout(self._ffn(s2), self.p, tr)
        a = PF.multi_head_attention(self.self_attn, src, pos, None, None, src_key_padding_mask, tr, pos_head=pos_head)
        # norm1 feeds the FFN (bf16 copy), norm2 feeds the next layer's attention / the decoder's
        # cross-attention (bf16(y) for values, bf16(y + pos) for queries and keys)
        src = PF.add_dropout_layernorm(a, src, self.norm1, self.p, tr, cast=True, x_exclusive=True)
        return PF.add_dropout_layernorm(self._ffn(src), src, self.norm2, self.p, tr, cast=True, cast_pos=pos, x_exclusive=True)


class TransformerEncoder(nn.Module):
    """transformer.py:118-158."""

    def __init__(self, d_model=256, nhead=8, dim_feedforward=2048, dropout=0.1, activation="relu",
                 normalize_before=False, num_layers=4):
        super().__init__()
        self.layers = nn.ModuleList([TransformerEncoderLayer(d_model, nhead, dim_feedforward, dropout, activation,
                                                             normalize_before) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = nn.LayerNorm(d_model) if normalize_before else None

    def forward(self, src, mask=None, src_key_padding_mask=None, pos=None, pos_head=None):
        out = src
        for layer in self.layers:
            out = layer(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask, pos=pos, pos_head=pos_head)
        if self.norm is not None:
            out = PF.add_dropout_layernorm(None, out, self.norm, 0.0, False)
        return out


class TransformerDecoderLayer(nn.Module):
    """transformer.py:286-404."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False):
        super().__init__()
        _check_activation(activation)
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.normalize_before = normalize_before
        self.p = dropout

    def _ffn(self, x):
        return PF.feed_forward(x, self.linear1, self.linear2, self.p, self.training)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos=None, query_pos=None, pos_head=None, memkv=None):
        """`memkv` (extension): this layer's slot of the memory keys / values projected for all layers at once by
        `TransformerDecoder.forward` (functional.memory_kv)."""
        assert tgt_mask is None and memory_mask is None and tgt_key_padding_mask is None
        tr = self.training
        if self.normalize_before:
            ln = (lambda x, n: PF.add_dropout_layernorm(None, x, n, 0.0, False))
            t2 = ln(tgt, self.norm1)
            tgt = tgt + PF.dropout(PF.multi_head_attention(self.self_attn, t2, query_pos, None, None, None, tr), self.p, tr)
            t2 = ln(tgt, self.norm2)
            tgt = tgt + PF.dropout(PF.multi_head_attention(self.multihead_attn, t2, query_pos, memory, pos,
                                                           memory_key_padding_mask, tr, mem_pos_head=pos_head), self.p, tr)
            t2 = ln(tgt, self.norm3)
            return tgt + PF.dropout(self._ffn(t2), self.p, tr)
        a = PF.multi_head_attention(self.self_attn, tgt, query_pos, None, None, None, tr)
        tgt = PF.add_dropout_layernorm(a, tgt, self.norm1, self.p, tr, cast_pos=query_pos, x_exclusive=True)  # -> cross-attention queries
        a = PF.multi_head_attention(self.multihead_attn, tgt, query_pos, memory, pos, memory_key_padding_mask, tr,
                                    mem_pos_head=pos_head, memkv=memkv)
        tgt = PF.add_dropout_layernorm(a, tgt, self.norm2, self.p, tr, cast=True, x_exclusive=True)  # -> FFN
        # -> next layer's self-attention
        return PF.add_dropout_layernorm(self._ffn(tgt), tgt, self.norm3, self.p, tr, cast=True, cast_pos=query_pos,
                                        x_exclusive=True)


class _FirstOfStack(torch.autograd.Function):
    """`torch.stack(intermediate)[0]` for a consumer that reads only the first decoder layer's output (ACT: act.py:262-270)
    without materialising the stack: returns intermediate[0]; in backward the other layers receive ONE shared zero tensor
    (they still run their backward, like the reference's select-of-stack does, but the 7-fold zero-filled stack, its
    copy and the seven strided-to-contiguous copies in front of the LayerNorm backward kernels are gone)."""

    @staticmethod
    def forward(ctx, *inter):
        ctx.n = len(inter)
        return inter[0].view_as(inter[0])

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        z = torch.zeros_like(g) if ctx.n > 1 else None
        return (g,) + (z,) * (ctx.n - 1)


class TransformerDecoder(nn.Module):
    """transformer.py:161-207.  `skip_dead_layers` (off by default = like-for-like with the
    reference) stops after the first layer when only intermediate [0] is consumed downstream."""

    def __init__(self, decoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        import copy

        self.layers = nn.ModuleList([copy.deepcopy(decoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate
        self.skip_dead_layers = False
        self.group_memory_kv = True  # A/B switch: False = every layer projects the memory itself (reference structure)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos=None, query_pos=None, pos_head=None, first_only=False):
        """`first_only` (extension): return intermediate[0] (Q, B, E) instead of the stack -- see _FirstOfStack."""
        out, inter = tgt, []
        ln = (lambda x: PF.add_dropout_layernorm(None, x, self.norm, 0.0, False))
        # the memory is the same for every layer: project it to all layers' keys / values in ONE launch
        # (post-LN layers only; the reference recomputes the projections inside each layer, transformer.py:317-346)
        live = self.layers[:1] if (self.skip_dead_layers and self.return_intermediate) else self.layers
        kv = None
        if self.group_memory_kv and not self.layers[0].normalize_before and len(live) > 0:
            kv = PF.memory_kv([l.multihead_attn for l in live], memory, pos, pos_head)
        for li, layer in enumerate(self.layers):
            out = layer(out, memory, memory_key_padding_mask=memory_key_padding_mask, pos=pos, query_pos=query_pos,
                        pos_head=pos_head, memkv=None if (kv is None or li >= len(live)) else (kv, li))
            if self.return_intermediate:
                inter.append(ln(out))
                if self.skip_dead_layers and li == 0:
                    return inter[0] if first_only else torch.stack(inter)
        if self.norm is not None:
            out = ln(out)
            if self.return_intermediate:
                inter[-1] = out
        if self.return_intermediate:
            if first_only:
                return _FirstOfStack.apply(*inter)
            return torch.stack(inter)
        return out.unsqueeze(0)


class Transformer(nn.Module):
    """transformer.py:16-115."""

    def __init__(self, d_model=512, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, activation="relu", normalize_before=False, return_intermediate_dec=False):
        super().__init__()
        self.encoder = TransformerEncoder(d_model=d_model, nhead=nhead, dim_feedforward=dim_feedforward,
                                          dropout=dropout, activation=activation, normalize_before=normalize_before,
                                          num_layers=num_encoder_layers)
        layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, activation, normalize_before)
        self.decoder = TransformerDecoder(layer, num_decoder_layers, nn.LayerNorm(d_model),
                                          return_intermediate=return_intermediate_dec)
        self._reset_parameters()
        self.d_model = d_model
        self.nhead = nhead
        # `src` / `pos_embed` may be (b, c, 1, n) views that carry the complete seq-first (S, B, E) buffers they were cut
        # from (attribute `_pcm_tokens`, see act.ACTPCD.forward_pcd_embed): the concatenations below are then skipped
        self.accepts_token_buffers = True

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, src, mask, query_embed, pos_embed, latent_input=None, proprio_input=None,
                additional_pos_embed=None, first_only=False):
        """`first_only` (extension, used by act.ACTPCD): return hs[0] as a (B, Q, E) view instead of the (n_layers, B, Q, E)
        stack the reference returns (transformer.py:106-115) -- every layer is still computed, forward and backward."""
        bs = src.shape[0]
        tok, ptk = getattr(src, "_pcm_tokens", None), getattr(pos_embed, "_pcm_tokens", None)
        if (tok is not None and ptk is not None and latent_input is not None and proprio_input is not None
                and additional_pos_embed is not None and getattr(ptk, "_pcm_add_pos", None) is additional_pos_embed
                and tok.shape == ptk.shape):
            E = tok.shape[2]
            prop = proprio_input.reshape(-1, bs, E)
            if 1 + prop.shape[0] + src.shape[-1] == tok.shape[0]:
                # token buffers written by the set-abstraction head / the sine-embedding kernel: only the latent / proprio
                # rows are missing (one small kernel), the learned positional rows get their gradient through `pos_head`
                src_tok = PF.fill_head_rows(tok, latent_input.reshape(bs, E), prop, ptk)
                PF.grad_boundary(src_tok, "transformer.encoder")
                pos_head = additional_pos_embed.unsqueeze(1) if additional_pos_embed.requires_grad else None
                query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
                tgt = torch.zeros_like(query_embed)
                with PF.stage("encoder x%d" % self.encoder.num_layers):
                    memory = self.encoder(src_tok, src_key_padding_mask=mask, pos=ptk, pos_head=pos_head)
                PF.grad_boundary(memory, "transformer.decoder")
                with PF.stage("decoder x%d" % self.decoder.num_layers):
                    hs = self.decoder(tgt, memory, memory_key_padding_mask=mask, pos=ptk, query_pos=query_embed, pos_head=pos_head,
                                      first_only=first_only and self.decoder.return_intermediate)
                if first_only and self.decoder.return_intermediate:
                    return hs.transpose(0, 1)
                return hs.transpose(1, 2)
        src = src.flatten(2).permute(2, 0, 1)
        pos_embed = pos_embed.flatten(2).permute(2, 0, 1)
        if pos_embed.shape[1] == 1:
            pos_embed = pos_embed.repeat(1, bs, 1)
        query_embed = query_embed.unsqueeze(1).repeat(1, bs, 1)
        # [learned rows ; sine embedding]: when the sine part is a constant (it always is on the
        # reference path: a function of the input coordinates) only the learned rows need a
        # gradient -- hand them over separately instead of differentiating through the concatenation
        pos_head = None
        if pos_embed.requires_grad or not additional_pos_embed.requires_grad:
            pos_embed = torch.cat([additional_pos_embed.unsqueeze(1).repeat(1, bs, 1), pos_embed], dim=0)
        else:
            pos_head = additional_pos_embed.unsqueeze(1)  # (n_add, 1, E), broadcast over the batch
            pos_embed = torch.cat([additional_pos_embed.detach().unsqueeze(1).repeat(1, bs, 1), pos_embed], dim=0)
        if latent_input.dim() == 2:
            addition_input = torch.stack([latent_input, proprio_input], dim=0)
        else:
            addition_input = torch.cat([latent_input, proprio_input], dim=0)
        src = torch.cat([addition_input, src], dim=0)
        PF.grad_boundary(src, "transformer.encoder")
        tgt = torch.zeros_like(query_embed)
        memory = self.encoder(src, src_key_padding_mask=mask, pos=pos_embed, pos_head=pos_head)
        PF.grad_boundary(memory, "transformer.decoder")
        hs = self.decoder(tgt, memory, memory_key_padding_mask=mask, pos=pos_embed, query_pos=query_embed,
                          pos_head=pos_head)
        if first_only:
            return hs.transpose(1, 2)[0]
        return hs.transpose(1, 2)
