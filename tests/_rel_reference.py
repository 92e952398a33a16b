"""torch restatement (TEST INFRASTRUCTURE) of the relative-positional self-attention the rel_att mode follows:
HF `Wav2Vec2ConformerSelfAttention.forward/_apply_relative_embeddings`
(transformers/models/wav2vec2_conformer/modeling_wav2vec2_conformer.py:447-565) with the packed in-projection
layout of nn.MultiheadAttention, so it can stand in for `self_attn` inside torchaudio's ConformerLayer
(conformer.py:194-200) and carry exactly our state_dict keys.  Pinned to the HF golden fixture by
tests/test_oracle_cpu.py::test_rel_torch_reference_matches_hf_golden; differentiable, any dtype.
"""
import math

import torch
from torch import nn


def rel_table(T: int, d: int) -> torch.Tensor:
    """HF Wav2Vec2ConformerRelPositionalEmbedding.extend_pe (:168-196): fp32, (2T-1, d), row r = position T-1-r."""
    pos = torch.arange(0, T, dtype=torch.int64).float().unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.int64).float() * -(math.log(10000.0) / d))
    pe_pos, pe_neg = torch.zeros(T, d), torch.zeros(T, d)
    pe_pos[:, 0::2], pe_pos[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
    pe_neg[:, 0::2], pe_neg[:, 1::2] = torch.sin(-1 * pos * div), torch.cos(-1 * pos * div)
    return torch.cat([torch.flip(pe_pos, [0]), pe_neg[1:]], 0)


def rel_shift(bd: torch.Tensor) -> torch.Tensor:
    """HF :553-559 (zero-pad, reshape, drop first row): out[..., i, j] = bd[..., i, T-1-i+j]."""
    B, H, T, P = bd.shape
    pad = torch.cat([torch.zeros(B, H, T, 1, dtype=bd.dtype, device=bd.device), bd], -1).view(B, H, P + 1, T)
    return pad[:, :, 1:].reshape(B, H, T, P)[..., : P // 2 + 1]


def rel_attention_core(qu, k, v, qv, pw, lengths, scale):
    """qu,k,v,qv: (B,H,T,dh); pw: (H, 2T-1, dh); lengths (B,) -> (B,H,T,dh)"""
    T = qu.shape[2]
    ac = qu @ k.transpose(-1, -2)
    bd = rel_shift(qv @ pw.transpose(-1, -2)[None])
    s = (ac + bd) * scale
    mask = torch.arange(T, device=qu.device)[None, :] >= lengths.to(qu.device)[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    return torch.softmax(s, -1) @ v


class RelMultiheadAttention(nn.Module):
    """Drop-in for `ConformerLayer.self_attn` (called with (T,B,D) tensors, key_padding_mask True = pad)."""

    def __init__(self, d: int, H: int):
        super().__init__()
        self.d, self.H = d, H
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        self.linear_pos = nn.Linear(d, d, bias=False)
        self.pos_bias_u = nn.Parameter(torch.zeros(H, d // H))
        self.pos_bias_v = nn.Parameter(torch.zeros(H, d // H))
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)

    def forward(self, query, key=None, value=None, key_padding_mask=None, need_weights=False, pe=None):
        T, B, d = query.shape
        H, dh = self.H, d // self.H
        x = query.transpose(0, 1)
        q, k, v = (torch.nn.functional.linear(x, self.in_proj_weight, self.in_proj_bias)
                   .view(B, T, 3, H, dh).unbind(2))
        if pe is None:
            pe = rel_table(T, d).to(device=x.device, dtype=x.dtype)
        pw = self.linear_pos(pe.view(2 * T - 1, d)).view(2 * T - 1, H, dh).transpose(0, 1)
        lengths = (~key_padding_mask).sum(1) if key_padding_mask is not None else torch.full((B,), T)
        o = rel_attention_core((q + self.pos_bias_u).transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                               (q + self.pos_bias_v).transpose(1, 2), pw, lengths, 1.0 / math.sqrt(dh))
        return self.out_proj(o.transpose(1, 2).reshape(B, T, d)).transpose(0, 1), None


def torchaudio_conformer_rel(d, H, f, L, k, dropout=0.0):
    """torchaudio.models.Conformer whose layers use RelMultiheadAttention: the fp64 reference of the rel_att mode."""
    from torchaudio.models import Conformer
    m = Conformer(d, H, f, L, k, dropout)
    for layer in m.conformer_layers:
        layer.self_attn = RelMultiheadAttention(d, H)
    return m
