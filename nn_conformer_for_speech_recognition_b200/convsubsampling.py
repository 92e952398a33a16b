"""Drop-in for the reference's `lib.convsubsampling.ConvSubSampling` (/root/reference/lib/convsubsampling.py:16-45).

Same constructor `ConvSubSampling(hp, in_size, out_nodes)`, same attributes (`hp`, `out_size`, `conv_sub_1`,
`conv_sub_2` -> identical state_dict keys and initialisation), same `forward(x (B,in,F,T)) -> (B,out,F',T')`.
The two nn.Conv2d members only HOLD the parameters; the arithmetic is the exact composition
conv_sub_2 o conv_sub_1 = Conv2d(in, out, k1+(k2-1)*s1, stride s1*s2) (nothing sits between the two convs,
convsubsampling.py:43-44) executed as a tcgen05 implicit GEMM (csrc/convsub.cu).  Gradients flow back to
conv_sub_1 / conv_sub_2 through the composition, so an optimizer updates the reference's own parameters.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib, ops


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class ConvSubSampling(nn.Module):
    def __init__(self, hp, in_size, out_nodes):
        super().__init__()
        self.hp = hp
        self.conv_sub_1 = nn.Conv2d(in_size, hp.conv_sub_1_nodes, hp.conv_sub_1_kernel, stride=hp.conv_sub_1_stride)
        self.conv_sub_2 = nn.Conv2d(hp.conv_sub_1_nodes, out_nodes, hp.conv_sub_2_kernel, stride=hp.conv_sub_2_stride)
        # out_size arithmetic exactly as convsubsampling.py:24-32
        conv_kers = [hp.conv_sub_1_kernel, hp.conv_sub_2_kernel]
        conv_strides = [hp.conv_sub_1_stride, hp.conv_sub_2_stride]
        convh, convw = hp.input_rows, hp.input_cols
        for i in range(len(conv_kers)):
            sh, sw = list(conv_strides[i])
            convh = (convh - conv_kers[i] + sh) // sh
            convw = (convw - conv_kers[i] + sw) // sw
        self.out_size = out_nodes * convh * convw

        s1h, s1w = _pair(hp.conv_sub_1_stride)
        s2h, s2w = _pair(hp.conv_sub_2_stride)
        if in_size != 1 or s1h != s1w or s2h != s2w:
            raise NotImplementedError("the sm_100a path covers the reference's use: in_size=1 (asrnn.py:27) and square strides")
        self.k1, self.k2, self.s1, self.s2 = hp.conv_sub_1_kernel, hp.conv_sub_2_kernel, s1h, s2h
        self.kc = self.k1 + (self.k2 - 1) * self.s1
        self.stride = self.s1 * self.s2
        self.KP = ((self.kc * self.kc + 63) // 64) * 64
        if self.KP > 256 or out_nodes > 256:
            raise NotImplementedError("composed kernel larger than 16x16 or more than 256 output channels")
        self.out_nodes = out_nodes
        self.C2p = ((out_nodes + 15) // 16) * 16          # UMMA N granularity; extra channels are zero and sliced off
        self._key = None
        self._ptrs = None
        self._composed = None

    # -- composed weights: static buffers, recomposed in place when a parameter changed -- and ALWAYS while the stream
    #    is being captured, so a replayed CUDA graph recomposes from the parameters the optimizer has just updated
    def composed(self):
        ps = (self.conv_sub_1.weight, self.conv_sub_1.bias, self.conv_sub_2.weight, self.conv_sub_2.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        ptrs = tuple(k[0] for k in key)
        if self._composed is None or ptrs != self._ptrs:
            dev = ps[0].device
            self._composed = (torch.zeros((self.C2p, self.KP), device=dev, dtype=torch.bfloat16),
                              torch.zeros(self.C2p, device=dev, dtype=torch.float32))
            self._ptrs, self._key = ptrs, None
        if key != self._key or torch.cuda.is_current_stream_capturing():
            ops.convsub_compose(*(p.detach() for p in ps), self.s1, self.KP, out=self._composed)   # rows >= out_nodes stay zero
            self._key = key
        return self._composed

    def forward_tokens(self, x: torch.Tensor, out_dtype=torch.bfloat16) -> torch.Tensor:
        """x (B,1,F,T) or (B,F,T) fp32 -> token-major (B, T', F', out_nodes); differentiable wrt the conv parameters."""
        _lib.require_device()
        if x.dim() == 4:
            if x.shape[1] != 1:
                raise ValueError("expected a single input channel")
            x = x[:, 0]
        if not x.is_cuda:
            raise _lib.B2CError("ConvSubSampling needs CUDA tensors; there is no CPU fallback")
        x = x.to(torch.float32).contiguous()
        ps = (self.conv_sub_1.weight, self.conv_sub_1.bias, self.conv_sub_2.weight, self.conv_sub_2.bias)
        if torch.is_grad_enabled() and any(p.requires_grad for p in ps):
            return _ConvSubFn.apply(x, self, out_dtype, *ps)
        Wc, bc = self.composed()
        return ops.convsub_fwd(x, Wc, bc, self.kc, self.stride, out_dtype)[..., :self.out_nodes]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Reference signature: (B,in,F,T) -> (B,out,F',T') fp32 (a permuted view of the token-major result)."""
        return self.forward_tokens(x, torch.float32).permute(0, 3, 2, 1)


class _ConvSubFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mod, out_dtype, W1, b1, W2, b2):
        Wc, bc = mod.composed()
        ctx.mod = mod
        ctx.save_for_backward(x, W1, b1, W2)
        return ops.convsub_fwd(x, Wc, bc, mod.kc, mod.stride, out_dtype)[..., :mod.out_nodes]

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        x, W1, b1, W2 = ctx.saved_tensors
        B, To, Fo, C2 = dout.shape
        Cw = ((C2 + 127) // 128) * 128                            # the wgrad kernel works on 128-channel slabs
        if Cw != C2:                                              # small / odd channel counts: zero-pad (rare path)
            pad = torch.zeros((B, To, Fo, Cw), device=dout.device, dtype=torch.bfloat16)
            pad[..., :C2] = dout
            dout = pad
        elif dout.dtype != torch.bfloat16 or not dout.is_contiguous():
            dout = dout.contiguous().to(torch.bfloat16)          # module-API path only; the fused pipeline hands bf16
        dWc = torch.zeros((Cw, mod.KP), device=x.device, dtype=torch.float32)
        dbc = torch.zeros(Cw, device=x.device, dtype=torch.float32)
        ops.convsub_wgrad(x, dout.view(B * To, Fo * Cw), dWc, mod.kc, mod.stride)
        ops.colsum(dout.view(-1, Cw), dbc)
        dWc, dbc = dWc[:C2].contiguous(), dbc[:C2].contiguous()
        dW1, db1 = torch.zeros_like(W1), torch.zeros_like(b1)
        dW2, db2 = torch.zeros_like(W2), torch.zeros(C2, device=x.device, dtype=torch.float32)
        ops.convsub_compose_bwd(dWc, dbc, W1.detach(), b1.detach(), W2.detach(), mod.s1, dW1, db1, dW2, db2)
        return None, None, None, dW1, db1, dW2, db2          # input gradient is never needed (log-mel input)
