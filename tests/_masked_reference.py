"""Differentiable fp64 restatement of ONE torchaudio ConformerLayer in train mode with EXPLICIT dropout masks.

TEST INFRASTRUCTURE.  torchaudio applies nn.Dropout at seven places per layer (conformer.py:74,106,108,152,153: after SiLU and after the
second Linear of each FFN, on the attention probabilities, after out_proj, after the conv module's second pointwise conv).
torch's Philox stream cannot be matched by the kernels, so the dropout-on parity test exports the masks
the kernels really drew (b2c_dropout_export_mask / the B2C_DROPPED_PREACT markers of the saved pre-activations) and this
function re-applies them as scale tensors (0 or 1/keep).  Everything else follows SURVEY.md Appendix A, checked against
torchaudio at p = 0 in tests/test_dropout_gpu.py::test_masked_reference_equals_torchaudio_without_dropout.
"""
import math

import torch
import torch.nn.functional as F


def layer_forward(x, P, lengths, masks, num_heads, bn_train=True, bn_eps=1e-5, pos=None):
    """x (B,T,d) fp64; P: dict of the layer's parameters/buffers (torchaudio state_dict keys without the layer prefix);
    masks: dict with optional scale tensors 'ffn1_h' (B,T,f), 'ffn1_o' (B,T,d), 'attn' (B,H,T,T), 'attn_o', 'conv_o',
    'ffn2_h', 'ffn2_o'.  Returns the layer output (B,T,d)."""
    B, T, d = x.shape
    H = num_heads
    dh = d // H
    one = torch.ones((), dtype=x.dtype)
    m = lambda k: masks.get(k, one)                                            # noqa: E731

    def ffn(z, pre):
        a = F.layer_norm(z, (d,), P[pre + ".sequential.0.weight"], P[pre + ".sequential.0.bias"])
        u = F.linear(a, P[pre + ".sequential.1.weight"], P[pre + ".sequential.1.bias"])
        h = F.silu(u) * m(pre + "_h")
        return F.linear(h, P[pre + ".sequential.4.weight"], P[pre + ".sequential.4.bias"]) * m(pre + "_o")

    x = x + 0.5 * ffn(x, "ffn1")                                               # conformer.py:185-187
    xa = x if pos is None else x + pos[:B * T].view(B, T, d)        # unused_lib/conformer.py:101: x + rel_pos_emb in front of the LN
    a = F.layer_norm(xa, (d,), P["self_attn_layer_norm.weight"], P["self_attn_layer_norm.bias"])
    qkv = F.linear(a, P["self_attn.in_proj_weight"], P["self_attn.in_proj_bias"])
    q, k, v = (t.view(B, T, H, dh).transpose(1, 2) for t in qkv.split(d, dim=-1))
    s = q @ k.transpose(-1, -2) / math.sqrt(dh)
    key_pad = torch.arange(T)[None, :] >= torch.as_tensor(lengths)[:, None]    # conformer.py:9-15 (True = padded key)
    s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1) * m("attn")
    o = (p @ v).transpose(1, 2).reshape(B, T, d)
    x = x + F.linear(o, P["self_attn.out_proj.weight"], P["self_attn.out_proj.bias"]) * m("attn_o")   # :192-202
    a = F.layer_norm(x, (d,), P["conv_module.layer_norm.weight"], P["conv_module.layer_norm.bias"])
    u = F.linear(a, P["conv_module.sequential.0.weight"][:, :, 0], P["conv_module.sequential.0.bias"])
    g = u[..., :d] * torch.sigmoid(u[..., d:])
    wdw = P["conv_module.sequential.2.weight"]                                  # (d,1,k)
    kk = wdw.shape[-1]
    y = F.conv1d(g.transpose(1, 2), wdw, P["conv_module.sequential.2.bias"], padding=(kk - 1) // 2, groups=d)   # (B,d,T)
    if bn_train:
        mu = y.mean(dim=(0, 2), keepdim=True)
        var = y.var(dim=(0, 2), unbiased=False, keepdim=True)
    else:
        mu = P["conv_module.sequential.3.running_mean"].view(1, d, 1)
        var = P["conv_module.sequential.3.running_var"].view(1, d, 1)
    yh = (y - mu) / torch.sqrt(var + bn_eps)
    c = F.silu(yh * P["conv_module.sequential.3.weight"].view(1, d, 1) + P["conv_module.sequential.3.bias"].view(1, d, 1))
    c = c.transpose(1, 2)
    x = x + F.linear(c, P["conv_module.sequential.5.weight"][:, :, 0], P["conv_module.sequential.5.bias"]) * m("conv_o")
    x = x + 0.5 * ffn(x, "ffn2")                                               # :207-209
    return F.layer_norm(x, (d,), P["final_layer_norm.weight"], P["final_layer_norm.bias"])                # :211
