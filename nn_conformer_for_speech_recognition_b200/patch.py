"""`patch_asrnn(model)`: swap the hot-path pieces of a reference `ASRNN` instance for the sm_100a drop-ins.

The seams are exactly the three the reference has: `model.conv_sub_sampling`
(asrnn.py:27,206), `model.conformers` (asrnn.py:29,214) and the four SpecAugment bound methods (asrnn.py:91-192).
Parameters are carried over through state_dict (keys and shapes are identical); nothing else of ASRNN is touched.
"""
from __future__ import annotations

import types

from . import specaugment
from .conformer import Conformer
from .convsubsampling import ConvSubSampling


def conformer_from_torchaudio(ref, rel_att: bool = False) -> Conformer:
    """Build our Conformer with the constructor arguments of a torchaudio.models.Conformer and copy its weights.
    rel_att=True (`hp.rel_att`, lib/hparams.py:102) adds the relative-positional parameters (fresh initialisation
    unless `ref` already carries `self_attn.linear_pos.weight` / `pos_bias_u` / `pos_bias_v`)."""
    layer = ref.conformer_layers[0]
    lin = layer.ffn1.sequential[1]
    new = Conformer(lin.in_features, layer.self_attn.num_heads, lin.out_features, len(ref.conformer_layers),
                    layer.conv_module.sequential[2].kernel_size[0], layer.self_attn_dropout.p, rel_att=rel_att)
    new.load_state_dict(ref.state_dict(), strict=not rel_att)
    p = next(ref.parameters())
    return new.to(device=p.device, dtype=p.dtype).train(ref.training)


def convsub_from_reference(ref) -> ConvSubSampling:
    new = ConvSubSampling(ref.hp, ref.conv_sub_1.in_channels, ref.conv_sub_2.out_channels)
    new.load_state_dict(ref.state_dict())
    p = next(ref.parameters())
    return new.to(device=p.device, dtype=p.dtype).train(ref.training)


def patch_asrnn(model, specaug_mode: str = "reference", encoder: bool = False, lstm: bool = False):
    """encoder=True also replaces the callers either side of the hot path: `ASRNN.encoder`
    (bridge GEMM + fused dropout + sync-free length filter, asrnn.py:193-221) and `ASRNN.projection_block`
    (Linear -> SiLU -> BatchNorm1d, asrnn.py:73-89) -- see asrnn_path.py.  lstm=True swaps `model.lstm` (asrnn.py:38,252)
    for `b2c.LSTM` (lstm.py); `final_fc` / `log_softmax` stay the reference's."""
    if specaug_mode not in ("reference", "paper"):
        raise ValueError("specaug_mode must be 'reference' or 'paper'")
    model.conv_sub_sampling = convsub_from_reference(model.conv_sub_sampling)
    model.conformers = conformer_from_torchaudio(model.conformers, bool(getattr(getattr(model, "hp", None), "rel_att", False)))
    model._b2c_specaug_mode = specaug_mode
    for name in ("SpecAugment", "time_warping", "frequency_masking", "time_masking"):
        setattr(model, name, types.MethodType(getattr(specaugment, name), model))
    if encoder:
        from . import asrnn_path
        model.encoder = types.MethodType(asrnn_path.encoder, model)
        model.projection_block = types.MethodType(asrnn_path.projection_block, model)
    if lstm:
        from .lstm import lstm_from_reference
        model.lstm = lstm_from_reference(model.lstm)
    return model
