"""Drop-in for `torchaudio.models.Conformer`, the Conformer stack the reference constructs at
`/root/reference/lib/standard/asrnn.py:29` and calls at `asrnn.py:214`.

Same constructor signature, same `forward(input (B,T,D), lengths (B,)) -> (output (B,T,D), lengths)`, same
attribute tree and therefore the same `state_dict` keys/shapes -- a reference checkpoint
loads here and vice versa, and `Runner.load_model`'s `'conformer' in key` filter (`runner.py:74`) keeps working.
The torch.nn submodules below are PARAMETER HOLDERS ONLY (they also make `torch.manual_seed(s); Conformer(...)`
reproduce torchaudio's initialisation bit for bit); their `forward` is never called.  All arithmetic runs in the
hand-written sm_100a kernels of libb200conformer.so (see ops.py / include/b200conformer.h):

    per layer, on a token-major fp32 residual stream x [B*T, d]         (torchaudio conformer.py:176-212)
      x += 0.5 * W4 SiLU(W1 LN(x))             LN kernel, GEMM(+bias+SiLU), GEMM(+bias, *0.5, +residual)
      x += Wo Attn(Wqkv LN(x))                 LN, GEMM, flash attention (key-length mask), GEMM(+residual)
      x += Wpw2 SiLU(BN(DW(GLU(Wpw1 LN(x)))))  LN, GEMM, fused GLU/depthwise/BN/SiLU chain, GEMM(+residual)
      x += 0.5 * FFN2(x);  x = LN(x)

bf16 operands, fp32 accumulation, fp32 residual stream / LayerNorm / softmax / BatchNorm statistics.
There is no CPU or eager-PyTorch fallback: without the built library or off sm_100 the forward raises.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib, ops

HEAD_PAD = 64          # the attention kernel's head dimension
# inference: fused LN -> GEMM -> SiLU -> GEMM -> 1/2-residual kernel (csrc/ffn_fused.cu).  Parity-green but measured SLOWER than
# the three-launch path at the benchmark shapes (105 vs 54 us per half-step at M = 23 872, d = 256, f = 1024: every 128-token
# tile re-streams the 1 MB of weights through one SM's L2 port), so it is opt-in: B2C_FFN_FUSED=1.
_FFN_FUSED = os.environ.get("B2C_FFN_FUSED", "0") == "1"
_LAYER_CALLS = os.environ.get("B2C_LAYER_CALLS", "1") != "0"      # 0: eval issues every launch from Python (debug / profiling)
# the final LayerNorm of layer i and the first LayerNorm of layer i+1 run as ONE kernel, forward and backward
# (b2c_layernorm_fwd_pair / _bwd_pair); B2C_LN_PAIR=0 keeps them separate (A/B timing, bit-identity test)
_LN_PAIR = os.environ.get("B2C_LN_PAIR", "1") != "0"


class _FeedForwardModule(nn.Module):
    """Parameter holder mirroring torchaudio `_FeedForwardModule` (conformer.py:91-119)."""

    def __init__(self, input_dim: int, hidden_dim: int, dropout: float = 0.0) -> None:
        super().__init__()
        self.sequential = nn.Sequential(
            nn.LayerNorm(input_dim),
            nn.Linear(input_dim, hidden_dim, bias=True),
            nn.SiLU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, input_dim, bias=True),
            nn.Dropout(dropout),
        )


class _ConvolutionModule(nn.Module):
    """Parameter holder mirroring torchaudio `_ConvolutionModule` (conformer.py:18-88)."""

    def __init__(self, input_dim: int, num_channels: int, depthwise_kernel_size: int, dropout: float = 0.0,
                 bias: bool = False, use_group_norm: bool = False) -> None:
        super().__init__()
        if (depthwise_kernel_size - 1) % 2 != 0:
            raise ValueError("depthwise_kernel_size must be odd to achieve 'SAME' padding.")
        if use_group_norm:
            raise NotImplementedError("use_group_norm=True is not on the reference's path (asrnn.py:29 uses the "
                                      "BatchNorm default); only BatchNorm1d is implemented")
        self.layer_norm = nn.LayerNorm(input_dim)
        self.sequential = nn.Sequential(
            nn.Conv1d(input_dim, 2 * num_channels, 1, stride=1, padding=0, bias=bias),
            nn.GLU(dim=1),
            nn.Conv1d(num_channels, num_channels, depthwise_kernel_size, stride=1,
                      padding=(depthwise_kernel_size - 1) // 2, groups=num_channels, bias=bias),
            nn.BatchNorm1d(num_channels),
            nn.SiLU(),
            nn.Conv1d(num_channels, input_dim, kernel_size=1, stride=1, padding=0, bias=bias),
            nn.Dropout(dropout),
        )


class ConformerLayer(nn.Module):
    """Parameter holder mirroring torchaudio `ConformerLayer` (conformer.py:122-212)."""

    def __init__(self, input_dim: int, ffn_dim: int, num_attention_heads: int, depthwise_conv_kernel_size: int,
                 dropout: float = 0.0, use_group_norm: bool = False, convolution_first: bool = False,
                 rel_att: bool = False) -> None:
        super().__init__()
        if convolution_first:
            raise NotImplementedError("convolution_first=True is not on the reference's path (asrnn.py:29)")
        self.ffn1 = _FeedForwardModule(input_dim, ffn_dim, dropout=dropout)
        self.self_attn_layer_norm = nn.LayerNorm(input_dim)
        self.self_attn = nn.MultiheadAttention(input_dim, num_attention_heads, dropout=dropout)
        if rel_att:
            # additive extension: HF Wav2Vec2ConformerSelfAttention's names and initialisation
            # (modeling_wav2vec2_conformer.py:441-445, 866-868) hung on the same `self_attn` module
            dh = input_dim // num_attention_heads
            self.self_attn.linear_pos = nn.Linear(input_dim, input_dim, bias=False)
            self.self_attn.pos_bias_u = nn.Parameter(torch.zeros(num_attention_heads, dh))
            self.self_attn.pos_bias_v = nn.Parameter(torch.zeros(num_attention_heads, dh))
            nn.init.xavier_uniform_(self.self_attn.pos_bias_u)
            nn.init.xavier_uniform_(self.self_attn.pos_bias_v)
        self.self_attn_dropout = nn.Dropout(dropout)
        self.conv_module = _ConvolutionModule(input_dim=input_dim, num_channels=input_dim,
                                              depthwise_kernel_size=depthwise_conv_kernel_size, dropout=dropout,
                                              bias=True, use_group_norm=use_group_norm)
        self.ffn2 = _FeedForwardModule(input_dim, ffn_dim, dropout=dropout)
        self.final_layer_norm = nn.LayerNorm(input_dim)
        self.convolution_first = convolution_first


# order of the per-layer parameters handed to the autograd function (and of the gradients it returns)
_LAYER_PARAMS = [
    "ffn1.sequential.0.weight", "ffn1.sequential.0.bias", "ffn1.sequential.1.weight", "ffn1.sequential.1.bias",
    "ffn1.sequential.4.weight", "ffn1.sequential.4.bias",
    "self_attn_layer_norm.weight", "self_attn_layer_norm.bias", "self_attn.in_proj_weight", "self_attn.in_proj_bias",
    "self_attn.out_proj.weight", "self_attn.out_proj.bias",
    "conv_module.layer_norm.weight", "conv_module.layer_norm.bias", "conv_module.sequential.0.weight",
    "conv_module.sequential.0.bias", "conv_module.sequential.2.weight", "conv_module.sequential.2.bias",
    "conv_module.sequential.3.weight", "conv_module.sequential.3.bias", "conv_module.sequential.5.weight",
    "conv_module.sequential.5.bias",
    "ffn2.sequential.0.weight", "ffn2.sequential.0.bias", "ffn2.sequential.1.weight", "ffn2.sequential.1.bias",
    "ffn2.sequential.4.weight", "ffn2.sequential.4.bias",
    "final_layer_norm.weight", "final_layer_norm.bias",
]
_REL_PARAMS = ["self_attn.linear_pos.weight", "self_attn.pos_bias_u", "self_attn.pos_bias_v"]
# GEMM weights that need a bf16 copy (everything else is consumed as fp32 straight from the Parameter)
_BF16_WEIGHTS = ["ffn1.sequential.1.weight", "ffn1.sequential.4.weight", "self_attn.in_proj_weight",
                 "self_attn.out_proj.weight", "conv_module.sequential.0.weight", "conv_module.sequential.5.weight",
                 "ffn2.sequential.1.weight", "ffn2.sequential.4.weight"]


def _get(module: nn.Module, dotted: str) -> torch.Tensor:
    obj = module
    for part in dotted.split("."):
        obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
    return obj


class _WeightCache:
    """bf16 copies of the GEMM weights in STATIC device buffers, refreshed IN PLACE by one multi-tensor cast launch.

    Refresh policy: eagerly, when a Parameter's version changed since the last refresh; ALWAYS while the current stream
    is being captured into a CUDA graph -- the refresh kernels are then part of the graph, so every replay re-derives the
    bf16 operands from the fp32 masters the optimizer has just updated (no Python runs at replay time, a host-side
    version check cannot)."""

    def __init__(self, owner: "Conformer"):
        self.owner = owner
        self.key = None
        self.ptrs = None
        self.bf16: List[dict] = []
        self.table = None
        self.max_n = 0
        self.stage: List[dict] = []       # padded fp32 staging buffers (head_dim < 64 / rel_att only)

    def _names(self):
        return _BF16_WEIGHTS + ["self_attn.in_proj_bias"] + (_REL_PARAMS if self.owner.rel_att else [])

    def _signature(self):
        sig = []
        for layer in self.owner.conformer_layers:
            for name in self._names():
                p = _get(layer, name)
                sig.append((p.data_ptr(), p._version))
        return tuple(sig)

    @property
    def padded(self) -> bool:
        return self.owner.head_dim != HEAD_PAD or self.owner.rel_att

    def _allocate(self, dev):
        """(Re)build the static buffers and the cast table; runs when the parameters' storage moved (.to(), first use)."""
        o = self.owner
        H, d = o.num_heads, o.input_dim
        nslab = 4 if o.rel_att else 3
        self.bf16, self.stage, entries = [], [], []
        bf = torch.bfloat16
        for layer in o.conformer_layers:
            out, stg = {}, {}
            for name in _BF16_WEIGHTS + (["self_attn.linear_pos.weight"] if o.rel_att else []):
                p = _get(layer, name)
                src = p
                if self.padded and name == "self_attn.in_proj_weight":
                    # heads zero-padded to 64; rel_att: a 4th slab repeats W_q -- the GEMM then emits [q+u | k | v | q+v]
                    src = stg[name] = torch.zeros(nslab * H * HEAD_PAD, d, device=dev, dtype=torch.float32)
                elif self.padded and name == "self_attn.out_proj.weight":
                    src = stg[name] = torch.zeros(d, H * HEAD_PAD, device=dev, dtype=torch.float32)
                elif name == "self_attn.linear_pos.weight":
                    src = stg[name] = torch.zeros(H * HEAD_PAD, d, device=dev, dtype=torch.float32)
                t = torch.empty(src.shape[:2] if src.dim() == 3 else src.shape, device=dev, dtype=bf)
                out[name] = t
                entries.append((src.data_ptr(), t.data_ptr(), src.numel()))
            if self.padded:
                out["in_bias"] = stg["in_bias"] = torch.zeros(nslab * H * HEAD_PAD, device=dev, dtype=torch.float32)
            else:
                out["in_bias"] = _get(layer, "self_attn.in_proj_bias")
            self.bf16.append(out)
            self.stage.append(stg)
        self.table = torch.tensor([v for e in entries for v in e], dtype=torch.int64).to(dev)
        self.ntab = len(entries)
        self.max_n = max(e[2] for e in entries)
        self._structs = None

    def _refresh(self):
        """fp32 masters -> bf16 operands, in place (capturable: device-side copies and ONE multi-tensor cast launch)."""
        o = self.owner
        if self.padded:
            H, dh, d = o.num_heads, o.head_dim, o.input_dim
            nslab = 4 if o.rel_att else 3
            with torch.no_grad():
                for layer, stg in zip(o.conformer_layers, self.stage):
                    w = stg["self_attn.in_proj_weight"].view(nslab, H, HEAD_PAD, d)
                    w[:3, :, :dh].copy_(_get(layer, "self_attn.in_proj_weight").view(3, H, dh, d))
                    stg["self_attn.out_proj.weight"].view(d, H, HEAD_PAD)[:, :, :dh].copy_(
                        _get(layer, "self_attn.out_proj.weight").view(d, H, dh))
                    b = stg["in_bias"].view(nslab, H, HEAD_PAD)
                    b[:3, :, :dh].copy_(_get(layer, "self_attn.in_proj_bias").view(3, H, dh))
                    if o.rel_att:
                        w[3].copy_(w[0])
                        b[3].copy_(b[0])
                        b[0, :, :dh].add_(_get(layer, "self_attn.pos_bias_u"))
                        b[3, :, :dh].add_(_get(layer, "self_attn.pos_bias_v"))
                        stg["self_attn.linear_pos.weight"].view(H, HEAD_PAD, d)[:, :dh].copy_(
                            _get(layer, "self_attn.linear_pos.weight").view(H, dh, d))
        _lib.call("b2c_multi_cast_f32_bf16", self.table.data_ptr(), self.ntab, self.max_n, _lib.stream_ptr())

    def get(self) -> List[dict]:
        sig = self._signature()
        ptrs = tuple(s[0] for s in sig)
        capturing = torch.cuda.is_current_stream_capturing()
        if ptrs != self.ptrs:
            self._allocate(next(self.owner.parameters()).device)
            self.ptrs, self.key = ptrs, None
        if capturing or sig != self.key:
            self._refresh()
            self.key = sig
        return self.bf16

    def layer_structs(self):
        """ctypes `b2c_layer_weights` per layer for the one-call-per-layer eval path (pointers are static)."""
        weights = self.get()
        if self._structs is None:
            out = []
            for layer, w in zip(self.owner.conformer_layers, weights):
                P = lambda name: _get(layer, name).data_ptr()                       # noqa: E731
                B = lambda name: w[name].data_ptr()                                 # noqa: E731
                lw = _lib.LayerWeights()
                for pre, dst in (("ffn1", lw.ffn1), ("ffn2", lw.ffn2)):
                    dst.ln_w, dst.ln_b = P(pre + ".sequential.0.weight"), P(pre + ".sequential.0.bias")
                    dst.w1, dst.b1 = B(pre + ".sequential.1.weight"), P(pre + ".sequential.1.bias")
                    dst.w2, dst.b2 = B(pre + ".sequential.4.weight"), P(pre + ".sequential.4.bias")
                lw.att_ln_w, lw.att_ln_b = P("self_attn_layer_norm.weight"), P("self_attn_layer_norm.bias")
                lw.in_w, lw.in_b = B("self_attn.in_proj_weight"), w["in_bias"].data_ptr()
                lw.out_w, lw.out_b = B("self_attn.out_proj.weight"), P("self_attn.out_proj.bias")
                lw.conv_ln_w, lw.conv_ln_b = P("conv_module.layer_norm.weight"), P("conv_module.layer_norm.bias")
                lw.pw1_w, lw.pw1_b = B("conv_module.sequential.0.weight"), P("conv_module.sequential.0.bias")
                lw.dw_w, lw.dw_b = P("conv_module.sequential.2.weight"), P("conv_module.sequential.2.bias")
                bn = layer.conv_module.sequential[3]
                lw.bn_w, lw.bn_b = bn.weight.data_ptr(), bn.bias.data_ptr()
                lw.bn_mean, lw.bn_var = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
                lw.pw2_w, lw.pw2_b = B("conv_module.sequential.5.weight"), P("conv_module.sequential.5.bias")
                lw.final_ln_w, lw.final_ln_b = P("final_layer_norm.weight"), P("final_layer_norm.bias")
                out.append(lw)
            self._structs = out
        return self._structs


class Conformer(nn.Module):
    r"""Conformer stack with torchaudio's interface (torchaudio/models/conformer.py:215-293).

    Args mirror `torchaudio.models.Conformer`: input_dim, num_heads, ffn_dim, num_layers,
    depthwise_conv_kernel_size, dropout=0.0, use_group_norm=False, convolution_first=False.
    The reference calls it positionally: `Conformer(512, 8, 512, 1, 33, 0.5)` (asrnn.py:29).
    """

    def __init__(self, input_dim: int, num_heads: int, ffn_dim: int, num_layers: int,
                 depthwise_conv_kernel_size: int, dropout: float = 0.0, use_group_norm: bool = False,
                 convolution_first: bool = False, rel_att: bool = False, pos_table: Optional[torch.Tensor] = None):
        super().__init__()
        # pos_table: the reference's own positional variant (unused_lib/conformer.py:101-103): a table from
        # relpos.RelativePositionalEmbeddings, one row per flattened token, ADDED to the attention branch's input in front
        # of its LayerNorm in every layer (a constant: not a parameter, not in the state_dict)
        self.pos_table = None if pos_table is None else pos_table.detach().to(torch.float32).contiguous()
        if input_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.conformer_layers = nn.ModuleList([
            ConformerLayer(input_dim, ffn_dim, num_heads, depthwise_conv_kernel_size, dropout=dropout,
                           use_group_norm=use_group_norm, convolution_first=convolution_first, rel_att=rel_att)
            for _ in range(num_layers)
        ])
        # rel_att (`hp.rel_att`, lib/hparams.py:102): Transformer-XL style relative-positional attention with HF's
        # rel_shift; adds self_attn.linear_pos.weight / pos_bias_u / pos_bias_v per layer.  Default off = torchaudio.
        self.rel_att = bool(rel_att)
        self.layer_params = _LAYER_PARAMS + (_REL_PARAMS if rel_att else [])
        self._pe_cache = {}
        self.input_dim, self.num_heads, self.ffn_dim = input_dim, num_heads, ffn_dim
        self.num_layers, self.kernel_size, self.dropout_p = num_layers, depthwise_conv_kernel_size, float(dropout)
        self.head_dim = input_dim // num_heads
        if self.head_dim > HEAD_PAD:
            raise NotImplementedError(f"head_dim {self.head_dim} > {HEAD_PAD} is not supported by the attention kernel "
                                      "(reference default, Conformer-S/M/L all have head_dim <= 64)")
        if input_dim % 8 != 0 or ffn_dim % 8 != 0 or input_dim > 1024:
            raise NotImplementedError("input_dim and ffn_dim must be multiples of 8, input_dim <= 1024")
        self._cache = _WeightCache(self)
        self._seed_counter = 0

    # ------------------------------------------------------------------------------------------------
    def forward(self, input: torch.Tensor, lengths: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        r"""input (B, T, input_dim); lengths (B,) valid frames per utterance. Returns (output (B,T,input_dim), lengths)."""
        _lib.require_device()
        if input.dim() != 3 or input.shape[-1] != self.input_dim:
            raise ValueError(f"expected input of shape (B, T, {self.input_dim}), got {tuple(input.shape)}")
        if not input.is_cuda:
            raise _lib.B2CError("Conformer.forward needs CUDA tensors; there is no CPU fallback")
        B, T, _ = input.shape
        if lengths.shape[0] != B:
            raise ValueError("lengths must have one entry per utterance")
        lengths_i32 = lengths.to(device=input.device, dtype=torch.int32)
        from .autograd import conformer_apply          # local import: autograd.py imports this module
        y = conformer_apply(self, input, lengths_i32)
        return y, lengths

    # ------------------------------------------------------------------------------------------------
    def rel_table(self, T: int, device) -> torch.Tensor:
        """HF Wav2Vec2ConformerRelPositionalEmbedding.extend_pe (modeling_wav2vec2_conformer.py:168-196): fp32 table,
        interleaved sin/cos, row r = relative position T-1-r.  Returned as bf16 [Rp, d] with Rp = 2T-1 rounded up to 8
        (zero rows), cached per T -- a constant, like a weight."""
        key = (T, str(device))
        if key not in self._pe_cache:
            d = self.input_dim
            pos = torch.arange(0, T, dtype=torch.int64).float().unsqueeze(1)
            div = torch.exp(torch.arange(0, d, 2, dtype=torch.int64).float() * -(math.log(10000.0) / d))
            pe_pos, pe_neg = torch.zeros(T, d), torch.zeros(T, d)
            pe_pos[:, 0::2], pe_pos[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
            pe_neg[:, 0::2], pe_neg[:, 1::2] = torch.sin(-1 * pos * div), torch.cos(-1 * pos * div)
            pe = torch.cat([torch.flip(pe_pos, [0]), pe_neg[1:]], 0)
            Rp = (2 * T - 1 + 7) // 8 * 8
            full = torch.zeros(Rp, d)
            full[:2 * T - 1] = pe
            self._pe_cache[key] = full.to(device=device, dtype=torch.bfloat16)
        return self._pe_cache[key]

    def next_seed(self) -> int:
        """Dropout seeds: torch's CPU generator picks the stream (so torch.manual_seed controls it), the counter
        keeps successive forward passes distinct."""
        self._seed_counter += 1
        base = int(torch.initial_seed()) & 0xFFFFFFFF
        return (base << 20) ^ (self._seed_counter * 0x9E3779B1)

    def extra_repr(self) -> str:
        return (f"input_dim={self.input_dim}, num_heads={self.num_heads}, ffn_dim={self.ffn_dim}, "
                f"num_layers={self.num_layers}, kernel={self.kernel_size}, dropout={self.dropout_p} [sm_100a kernels]")


# ----------------------------------------------------------------------------------------------------
# the layer pipeline (shared by eval forward and the autograd function's forward)
# ----------------------------------------------------------------------------------------------------
def run_layers(model: Conformer, x: torch.Tensor, lengths_i32: torch.Tensor, B: int, T: int, train_state: Optional[list],
               bn_train: bool, drop_p: float, seed: int, force_two_pass: bool = False) -> torch.Tensor:
    """x: fp32 [B*T, d] residual stream (overwritten in eval mode). `train_state` (a list) receives what the
    backward pass needs; None = inference (buffers are reused across layers, nothing is kept)."""
    d, f, H, k = model.input_dim, model.ffn_dim, model.num_heads, model.kernel_size
    M = B * T
    Hd = H * HEAD_PAD
    scale = 1.0 / math.sqrt(model.head_dim)
    weights = model._cache.get()
    dev = x.device
    save = train_state is not None
    bf = torch.bfloat16
    if not save and not bn_train and not force_two_pass and not model.rel_att and model.pos_table is None and _LAYER_CALLS:
        # eval: ONE C call per layer (b2c_conformer_layer_eval enqueues the same 20 launches): the per-launch Python /
        # ctypes cost otherwise exceeds the GPU time of a layer in the eager inference / NST sweep path
        import ctypes
        structs = model._cache.layer_structs()
        bn0 = model.conformer_layers[0].conv_module.sequential[3]
        nbytes = _lib.load().b2c_conformer_layer_eval_workspace(B, T, d, f, H)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        for lw in structs:
            _lib.call("b2c_conformer_layer_eval", x.data_ptr(), ctypes.byref(lw), lengths_i32.data_ptr(), B, T, d, f, H,
                      model.head_dim, k, bn0.eps, ws.data_ptr(), nbytes, _lib.stream_ptr())
        return x
    rel = model.rel_att
    nq = 4 if rel else 3                               # rel_att: the in-projection also emits q + pos_bias_v
    pe = model.rel_table(T, dev) if rel else None
    if not save:       # inference workspaces, reused by every layer (stay L2-resident between kernels)
        a_buf = torch.empty((M, d), device=dev, dtype=bf)
        h_buf = torch.empty((M, f), device=dev, dtype=bf)
        qkv_buf = torch.empty((M, nq * Hd), device=dev, dtype=bf)
        u_buf = torch.empty((M, 2 * d), device=dev, dtype=bf)

    # per-layer BatchNorm scratch (fp64 sum / sum of squares) zeroed by ONE fill, counters bumped by ONE multi-tensor add:
    # 2 launches per forward instead of 2 per layer (each ~2-3 us of an otherwise idle GPU)
    stats_all = torch.zeros((model.num_layers, 2 * d), device=dev, dtype=torch.float64) if (bn_train or force_two_pass) else None
    if bn_train:
        torch._foreach_add_([l.conv_module.sequential[3].num_batches_tracked for l in model.conformer_layers], 1)

    pending_ln = []        # (a, mean, rstd) of the next layer's first LayerNorm when the pair kernel has already produced it
    for li, layer in enumerate(model.conformer_layers):
        w = weights[li]
        P = lambda name: _get(layer, name)                                      # noqa: E731
        sd = (seed + li * 16) if drop_p > 0 else 0
        st = {} if save else None

        def ffn(x_in, pre, site):
            ln_w, ln_b = P(pre + ".sequential.0.weight"), P(pre + ".sequential.0.bias")
            # the first feed-forward module's LayerNorm may already have run, fused with the previous layer's final
            # LayerNorm (b2c_layernorm_fwd_pair at the bottom of this loop)
            done = pending_ln.pop() if (pre == "ffn1" and pending_ln) else None
            if save:
                a, mean, rstd = done if done is not None else ops.layernorm_fwd(x_in, ln_w, ln_b, save_stats=True)
                u = torch.empty((M, f), device=dev, dtype=bf)
                h = ops.gemm(a, w[pre + ".sequential.1.weight"], M=M, N=f, K=d, bias=P(pre + ".sequential.1.bias"),
                             act=ops.ACT_SILU, out2=u, drop_p=drop_p, drop_seed=sd + site)
                x_out = ops.gemm(h, w[pre + ".sequential.4.weight"], M=M, N=d, K=f, bias=P(pre + ".sequential.4.bias"),
                                 resid=x_in, out_dtype=torch.float32, alpha=0.5, drop_p=drop_p, drop_seed=sd + site + 1)
                st[pre] = dict(x=x_in, a=a, mean=mean, rstd=rstd, u=u, h=h)
                return x_out
            if _FFN_FUSED and ops.ffn_fused_ok(d, f):      # one kernel: the [M, f] hidden activation never reaches HBM
                return ops.ffn_fused_eval(x_in, ln_w, ln_b, w[pre + ".sequential.1.weight"], P(pre + ".sequential.1.bias"),
                                          w[pre + ".sequential.4.weight"], P(pre + ".sequential.4.bias"), out=x_in)
            ops.layernorm_fwd(x_in, ln_w, ln_b, out=a_buf)
            ops.gemm(a_buf, w[pre + ".sequential.1.weight"], M=M, N=f, K=d, bias=P(pre + ".sequential.1.bias"),
                     act=ops.ACT_SILU, out=h_buf)
            ops.gemm(h_buf, w[pre + ".sequential.4.weight"], M=M, N=d, K=f, bias=P(pre + ".sequential.4.bias"),
                     resid=x_in, out=x_in, alpha=0.5)
            return x_in

        # ---- 1. x += 0.5 * FFN1(x)                                                   conformer.py:185-187
        with _lib.nvtx(f"layer{li}.ffn1"):
            x = ffn(x, "ffn1", 0)
        _nv = _lib.nvtx(f"layer{li}.mhsa")
        _nv.__enter__()

        # ---- 2. x += out_proj(Attn(in_proj(LN(x))))                                  conformer.py:192-202
        ln_w, ln_b = P("self_attn_layer_norm.weight"), P("self_attn_layer_norm.bias")
        pw = None
        if rel:            # Pw = linear_pos(PE): [Rp, H*64] projected relative embeddings of this layer
            pw = ops.gemm(pe, w["self_attn.linear_pos.weight"], M=pe.shape[0], N=Hd, K=d)
        x_att = x                              # input of the attention LayerNorm (the residual below stays x)
        if model.pos_table is not None:        # x + E, unused_lib/conformer.py:101
            E = model.pos_table
            if E.shape[-1] != d or E.shape[0] < M:
                raise ValueError(f"pos_table must be (>= {M}, {d}); got {tuple(E.shape)}")
            if E.device != dev:
                E = model.pos_table = E.to(dev)
            x_att = torch.empty_like(x)
            _lib.call("b2c_add_f32", x.data_ptr(), E.data_ptr(), x_att.data_ptr(), M * d, _lib.stream_ptr())
        if save:
            a, mean, rstd = ops.layernorm_fwd(x_att, ln_w, ln_b, save_stats=True)
            qkv = ops.gemm(a, w["self_attn.in_proj_weight"], M=M, N=nq * Hd, K=d, bias=w["in_bias"])
            if rel:
                att, lse = ops.attention_rel_fwd(qkv, pw, lengths_i32, B, T, H, scale, save_lse=True, drop_p=drop_p,
                                                 drop_seed=sd + 4)
            else:
                att, lse = ops.attention_fwd(qkv, lengths_i32, B, T, H, scale, save_lse=True, drop_p=drop_p, drop_seed=sd + 4)
            x_out = ops.gemm(att, w["self_attn.out_proj.weight"], M=M, N=d, K=Hd, bias=P("self_attn.out_proj.bias"),
                             resid=x, out_dtype=torch.float32, drop_p=drop_p, drop_seed=sd + 5)
            st["attn"] = dict(x=x_att, a=a, mean=mean, rstd=rstd, qkv=qkv, att=att, lse=lse, pw=pw)
            x = x_out
        else:
            ops.layernorm_fwd(x_att, ln_w, ln_b, out=a_buf)
            ops.gemm(a_buf, w["self_attn.in_proj_weight"], M=M, N=nq * Hd, K=d, bias=w["in_bias"], out=qkv_buf)
            att = (ops.attention_rel_fwd(qkv_buf, pw, lengths_i32, B, T, H, scale) if rel
                   else ops.attention_fwd(qkv_buf, lengths_i32, B, T, H, scale))
            ops.gemm(att, w["self_attn.out_proj.weight"], M=M, N=d, K=Hd, bias=P("self_attn.out_proj.bias"),
                     resid=x, out=x)

        _nv.__exit__()
        # ---- 3. x += conv_module(x)                                                  conformer.py:168-174,77-88
        _nv = _lib.nvtx(f"layer{li}.conv_module")
        _nv.__enter__()
        bn = layer.conv_module.sequential[3]
        ln_w, ln_b = P("conv_module.layer_norm.weight"), P("conv_module.layer_norm.bias")
        wdw = P("conv_module.sequential.2.weight").view(d, k)
        bdw = P("conv_module.sequential.2.bias")
        if save:
            a, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, save_stats=True)
            u = ops.gemm(a, w["conv_module.sequential.0.weight"], M=M, N=2 * d, K=d, bias=P("conv_module.sequential.0.bias"))
        else:
            ops.layernorm_fwd(x, ln_w, ln_b, out=a_buf)
            u = ops.gemm(a_buf, w["conv_module.sequential.0.weight"], M=M, N=2 * d, K=d,
                         bias=P("conv_module.sequential.0.bias"), out=u_buf)
        if bn_train:
            y, stats = ops.convmod_dw_fwd(u, wdw, bdw, B, T, stats=stats_all[li])
            c, bmean, brstd = ops.convmod_bn_finalize_apply(y, stats, M, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                                            bn.eps, bn.momentum if bn.momentum is not None else 0.1)
        elif force_two_pass:     # eval-mode BatchNorm under autograd: keep y, normalise with the running statistics
            y, _ = ops.convmod_dw_fwd(u, wdw, bdw, B, T, stats=stats_all[li])
            bmean, brstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
            c = ops.convmod_bn_apply(y, bmean, brstd, bn.weight, bn.bias)
        else:
            c = ops.convmod_eval(u, wdw, bdw, bn.weight, bn.bias, bn.running_mean, bn.running_var, B, T, bn.eps)
            y = bmean = brstd = None
        if save:
            x_out = ops.gemm(c, w["conv_module.sequential.5.weight"], M=M, N=d, K=d, bias=P("conv_module.sequential.5.bias"),
                             resid=x, out_dtype=torch.float32, drop_p=drop_p, drop_seed=sd + 8)
            st["conv"] = dict(x=x, a=a, mean=mean, rstd=rstd, u=u, y=y, bmean=bmean, brstd=brstd, c=c, bn_train=bn_train)
            x = x_out
        else:
            ops.gemm(c, w["conv_module.sequential.5.weight"], M=M, N=d, K=d, bias=P("conv_module.sequential.5.bias"),
                     resid=x, out=x)

        _nv.__exit__()
        # ---- 4. x += 0.5 * FFN2(x)                                                   conformer.py:206-209
        with _lib.nvtx(f"layer{li}.ffn2"):
            x = ffn(x, "ffn2", 12)

        # ---- 5. x = LN(x)                                                            conformer.py:211
        # ... fused with the LayerNorm that follows it immediately, the one of the NEXT layer's first feed-forward module
        # (conformer.py:103,185): one pass over the rows instead of two (b2c_layernorm_fwd_pair)
        ln_w, ln_b = P("final_layer_norm.weight"), P("final_layer_norm.bias")
        nxt = model.conformer_layers[li + 1] if (_LN_PAIR and save and li + 1 < model.num_layers and d <= 512) else None
        if save:
            if nxt is not None:
                x_out, mean, rstd, a_n, mean_n, rstd_n = ops.layernorm_fwd_pair(
                    x, ln_w, ln_b, _get(nxt, "ffn1.sequential.0.weight"), _get(nxt, "ffn1.sequential.0.bias"), save_stats=True)
                pending_ln.append((a_n, mean_n, rstd_n))
            else:
                x_out, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, out_dtype=torch.float32, save_stats=True)
            st["final"] = dict(x=x, mean=mean, rstd=rstd)
            train_state.append(st)
            x = x_out
        else:              # inference keeps the launch sequence of b2c_conformer_layer_eval (bit-identical paths)
            x = ops.layernorm_fwd(x, ln_w, ln_b, out_dtype=torch.float32)
    return x
