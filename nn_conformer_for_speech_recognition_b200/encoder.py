"""`ConformerEncoder`: the reference's `ASRNN.encoder` hot path (asrnn.py:193-221) assembled from the drop-ins:

    [SpecAugment]  ->  ConvSubSampling  ->  bridge  ->  Conformer stack

The reference's bridge is `flatten -> Linear(out_size, 512*max_len) -> view` (asrnn.py:207-209), whose parameter
count grows with T*T' (28 G parameters at 10 s / 100 fps) and cannot be instantiated at the benchmark shapes;
as SURVEY.md section 8 prescribes, this encoder uses a per-frame `Linear(out_nodes*F', d)` on the token-major
conv-subsampling output instead (same position in the pipeline, same GEMM kernel).  At the notebook's 40x40 shapes
use `patch_asrnn` on the real ASRNN -- that keeps the reference bridge untouched.

Everything between the log-mel input and the encoder output runs in libb200conformer.so; torch supplies memory,
streams and the autograd graph nodes.
"""
from __future__ import annotations

import random
from typing import Optional

import torch
from torch import nn

from . import _lib, ops, specaugment
from .conformer import Conformer
from .convsubsampling import ConvSubSampling


class _BridgeFn(torch.autograd.Function):
    """y[M,d] = dropout(a[M,K] W^T + b) on the tcgen05 GEMM; a is the bf16 token-major conv-subsampling output.  The
    dropout is the reference's `self.dropout(x)` between bridge and Conformer stack (asrnn.py:210), fused as the GEMM's
    epilogue; the backward regenerates the same mask (same seed, same per-call epoch word) while casting dy."""

    @staticmethod
    def forward(ctx, a, weight, bias, w_bf16, drop_p, drop_seed):
        M, K = a.shape
        ctx.save_for_backward(a, w_bf16)
        ctx.shapes = (M, K, weight.shape[0])
        ctx.drop = (drop_p, drop_seed)
        ctx.epoch = _lib.epoch_snapshot(a.device) if drop_p > 0.0 else None
        with _lib.epoch_scope(ctx.epoch):
            return ops.gemm(a, w_bf16, M=M, N=weight.shape[0], K=K, bias=bias, out_dtype=torch.float32, drop_p=drop_p,
                            drop_seed=drop_seed)

    @staticmethod
    def backward(ctx, dy):
        a, w_bf16 = ctx.saved_tensors
        M, K, N = ctx.shapes
        drop_p, drop_seed = ctx.drop
        db = torch.zeros(N, device=dy.device, dtype=torch.float32)
        with _lib.epoch_scope(ctx.epoch):
            dyb = ops.scale_cast(dy.contiguous().view(M, N), 1.0, drop_p, drop_seed, colsum=db)
        dW = torch.zeros((N, K), device=dy.device, dtype=torch.float32)
        ops.gemm(dyb, a, M=N, N=K, K=M, a_mn=True, b_mn=True, out=dW, accumulate=True,
                 split_k=max(1, min(16, ((M + 63) // 64) // 4)))
        da = ops.gemm(dyb, w_bf16, M=M, N=K, K=N, b_mn=True) if ctx.needs_input_grad[0] else None
        return da, dW, db, None, None, None


class _LinearFn(torch.autograd.Function):
    """y[M,N] (fp32) = x[M,K] (fp32) W^T + b: bf16 operands on the tcgen05 GEMM, fp32 accumulation and output."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.detach().reshape(-1, K).to(torch.float32).contiguous()
        M = x2.shape[0]
        a = ops.scale_cast(x2, 1.0)
        w_bf16 = ops.cast_bf16(weight.detach().contiguous())
        ctx.save_for_backward(a, w_bf16)
        ctx.shape = x.shape
        y = ops.gemm(a, w_bf16, M=M, N=N, K=K, bias=bias, out_dtype=torch.float32)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        a, w_bf16 = ctx.saved_tensors
        M, K = a.shape
        N = w_bf16.shape[0]
        db = torch.zeros(N, device=dy.device, dtype=torch.float32)
        dyb = ops.scale_cast(dy.reshape(M, N).to(torch.float32).contiguous(), 1.0, colsum=db)
        dW = torch.zeros((N, K), device=dy.device, dtype=torch.float32)
        ops.gemm(dyb, a, M=N, N=K, K=M, a_mn=True, b_mn=True, out=dW, accumulate=True,
                 split_k=max(1, min(32, ((M + 63) // 64) // 4)))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dyb, w_bf16, M=M, N=K, K=N, b_mn=True, out_dtype=torch.float32).view(ctx.shape)
        return dx, dW, db


class Linear(nn.Linear):
    """`torch.nn.Linear` (same parameters / state_dict) whose forward and backward run on the tcgen05 GEMM kernel --
    the building block for the pieces either side of the encoder (projection block, CTC head,
    `asrnn.py:73-89,254`).  in_features and out_features must be multiples of 8."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.bias is None:
            raise NotImplementedError("bias=False is not wired (every Linear on the reference path has a bias)")
        return _LinearFn.apply(x, self.weight, self.bias)


class ConformerEncoder(nn.Module):
    def __init__(self, hp, input_dim: int, num_heads: int, ffn_dim: int, num_layers: int,
                 depthwise_conv_kernel_size: int, dropout: float = 0.0, conv_out_nodes: int = 128,
                 specaug_mode: str = "reference", bridge_dropout: Optional[float] = None):
        super().__init__()
        self.hp = hp
        # `self.dropout(x)` between bridge and Conformer stack (asrnn.py:210): the reference uses hp.dropout for both
        self.bridge_dropout_p = float(dropout if bridge_dropout is None else bridge_dropout)
        self.conv_sub_sampling = ConvSubSampling(hp, 1, conv_out_nodes)
        cs = self.conv_sub_sampling
        self.freq_out = (hp.input_rows - cs.kc) // cs.stride + 1
        self.bridge = nn.Linear(conv_out_nodes * self.freq_out, input_dim)
        self.conformers = Conformer(input_dim, num_heads, ffn_dim, num_layers, depthwise_conv_kernel_size, dropout)
        self._b2c_specaug_mode = specaug_mode
        self._bridge_key = None
        self._bridge_bf16 = None

    def output_lengths(self, input_lens: torch.Tensor) -> torch.Tensor:
        """Frames after the two valid convolutions: ((n - k1)//s1 + 1 - k2)//s2 + 1, clamped at 0."""
        cs = self.conv_sub_sampling
        l1 = torch.div(input_lens - cs.k1, cs.s1, rounding_mode="floor") + 1
        l2 = torch.div(l1 - cs.k2, cs.s2, rounding_mode="floor") + 1
        return l2.clamp(min=0)

    def _bridge_weight(self):
        """bf16 bridge weight in the kernel's token-major feature order, in a STATIC buffer.  Refreshed when the Parameter
        changed and always while the stream is being captured (a replayed graph must re-derive it from the fp32 master
        the optimizer updates -- same policy as conformer._WeightCache)."""
        w = self.bridge.weight
        key = (w.data_ptr(), w._version)
        if self._bridge_bf16 is None or self._bridge_bf16.device != w.device or key[0] != (self._bridge_key or (None,))[0]:
            self._bridge_f32 = torch.empty(w.shape, device=w.device, dtype=torch.float32)
            self._bridge_bf16 = torch.empty(w.shape, device=w.device, dtype=torch.bfloat16)
            self._bridge_key = None
        if key != self._bridge_key or torch.cuda.is_current_stream_capturing():
            # the module stores the weight with the reference's NCHW flatten order (feature = c*F' + f'); the kernel's
            # token-major rows are ordered f'*C + c: permuting copy + cast (weight prep, two launches)
            C, Fo = self.conv_sub_sampling.out_nodes, self.freq_out
            with torch.no_grad():
                self._bridge_f32.view(-1, Fo, C).copy_(w.detach().view(-1, C, Fo).permute(0, 2, 1))
            ops.cast_bf16(self._bridge_f32, self._bridge_bf16)
            self._bridge_key = key
        return self._bridge_bf16

    def forward(self, x: torch.Tensor, input_lens: torch.Tensor, SpecAugment: bool = False,
                lens_host: Optional[list] = None, specaug_plan: "Optional[specaugment.SpecAugPlan]" = None):
        """x (B,1,F,T) fp32 log-mel on the GPU, input_lens (B,) frames.  Returns (encoded (B,T',d) fp32, lengths (B,)).
        `specaug_plan`: pre-drawn parameters (SpecAugPlan.draw was called on the host) -- the CUDA-graph friendly form."""
        if SpecAugment and specaug_plan is not None:
            x = specaug_plan.apply(x)
        elif SpecAugment:
            x = specaugment.SpecAugment(self, x, lens_host if lens_host is not None else input_lens)
        with _lib.nvtx("encoder.conv_sub_sampling"):
            tok = self.conv_sub_sampling.forward_tokens(x)              # (B, T', F', C) bf16
        B, To, Fo, C = tok.shape
        w_bf16 = self._bridge_weight()
        drop_p = self.bridge_dropout_p if self.training else 0.0
        drop_seed = (self.conformers.next_seed() ^ 0x5DEECE66D) if drop_p > 0.0 else 0
        if torch.is_grad_enabled() and (self.bridge.weight.requires_grad or tok.requires_grad):
            wp_view = self.bridge.weight.view(-1, C, Fo).permute(0, 2, 1).reshape(self.bridge.weight.shape[0], Fo * C)
            h = _BridgeFn.apply(tok.view(B * To, Fo * C), wp_view, self.bridge.bias, w_bf16, drop_p, drop_seed)
        else:
            with _lib.epoch_scope(_lib.epoch_snapshot(tok.device) if drop_p > 0.0 else None):
                h = ops.gemm(tok.view(B * To, Fo * C), w_bf16, M=B * To, N=self.bridge.weight.shape[0], K=Fo * C,
                             bias=self.bridge.bias, out_dtype=torch.float32, drop_p=drop_p, drop_seed=drop_seed)
        out_lens = self.output_lengths(input_lens)
        y, _ = self.conformers(h.view(B, To, -1), out_lens)
        return y, out_lens
