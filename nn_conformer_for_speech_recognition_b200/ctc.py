"""Fused log-softmax + CTC loss on the GPU (csrc/ctc.cu) -- the step that follows the encoder in the reference's training
loop: `nn.LogSoftmax(-1)` (/root/reference/lib/standard/asrnn.py:44,254-256) and
`CTCLoss(blank=hp.blank_idx, zero_infinity=True)` (/root/reference/lib/standard/runner.py:35,143).

Two entry points:
    ctc_loss_from_logits(logits (B,T,V), targets (B,S), input_lengths, target_lengths, blank, zero_infinity)
        the fused form: raw logits in, scalar mean loss out, gradient w.r.t. the logits in ONE extra pass.
    CTCLoss(blank, zero_infinity)(log_probs (T,B,V), targets, input_lengths, target_lengths)
        drop-in for the reference's `self.loss(logits.transpose(0,1), target, output_lengths, target_lens)`
        (runner.py:143).  log_softmax is idempotent, so feeding already-normalised log-probabilities gives the same loss,
        and the gradient it returns, exp(log_probs) - posterior, is exactly what torch's CTC backward returns.
No CPU fallback: CUDA tensors only.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib


def _i32(v, device):
    if isinstance(v, torch.Tensor):
        return v.to(device=device, dtype=torch.int32).contiguous()
    return torch.tensor([int(a) for a in v], dtype=torch.int32).to(device)


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, in_lens, tgt_lens, blank, zero_infinity):
        B, T, V = logits.shape
        S = targets.shape[1]
        x = logits.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        nbytes = _lib.load().b2c_ctc_workspace(B, T, S)
        ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        nll = torch.empty(B, device=x.device, dtype=torch.float32)
        ctx.args = (B, T, V, S, int(blank), int(bool(zero_infinity)))
        _lib.call("b2c_ctc_loss", x.data_ptr(), targets.data_ptr(), in_lens.data_ptr(), tgt_lens.data_ptr(), B, T, V, S,
                  int(blank), int(bool(zero_infinity)), 1, None, loss.data_ptr(), nll.data_ptr(), None, ws.data_ptr(), nbytes,
                  _lib.stream_ptr(), nbytes=4.0 * B * T * V)
        ctx.save_for_backward(x, targets, in_lens, tgt_lens, ws, loss, nll)
        ctx.in_dtype = logits.dtype
        ctx.mark_non_differentiable(nll)
        return loss.view(()), nll

    @staticmethod
    def backward(ctx, g, _g_nll):
        x, targets, in_lens, tgt_lens, ws, loss, nll = ctx.saved_tensors
        B, T, V, S, blank, zi = ctx.args
        g = g.detach().to(torch.float32).contiguous()
        grad = torch.empty_like(x)
        _lib.call("b2c_ctc_loss", x.data_ptr(), targets.data_ptr(), in_lens.data_ptr(), tgt_lens.data_ptr(), B, T, V, S,
                  blank, zi, 2, g.data_ptr(), loss.data_ptr(), nll.data_ptr(), grad.data_ptr(), ws.data_ptr(), ws.numel(),
                  _lib.stream_ptr(), nbytes=8.0 * B * T * V)
        return grad.to(ctx.in_dtype), None, None, None, None, None


def ctc_loss_from_logits(logits: torch.Tensor, targets: torch.Tensor, input_lengths, target_lengths, blank: int = 0,
                         zero_infinity: bool = True, return_nll: bool = False):
    """logits (B,T,V) CUDA; targets (B,S) integer class indices (padded; only the first target_lengths[b] count).
    Returns the scalar loss of `CTCLoss(reduction='mean')` on log_softmax(logits) (and the per-utterance nll on request)."""
    _lib.require_device()
    if logits.dim() != 3:
        raise ValueError(f"expected logits of shape (B, T, V), got {tuple(logits.shape)}")
    if not logits.is_cuda:
        raise _lib.B2CError("ctc_loss_from_logits needs CUDA tensors; there is no CPU fallback")
    if targets.dim() != 2 or targets.shape[0] != logits.shape[0]:
        raise ValueError("targets must be (B, S): the padded 2-D form the reference uses (runner.py:136)")
    dev = logits.device
    tg = _i32(targets, dev)
    il, tl = _i32(input_lengths, dev), _i32(target_lengths, dev)
    if il.numel() != logits.shape[0] or tl.numel() != logits.shape[0]:
        raise ValueError("input_lengths and target_lengths need one entry per utterance")
    loss, nll = _CTCFn.apply(logits, tg, il, tl, blank, zero_infinity)
    return (loss, nll) if return_nll else loss


class CTCLoss(nn.Module):
    """`torch.nn.CTCLoss(blank, reduction='mean', zero_infinity)` for (T,B,V) log-probabilities (runner.py:35,143)."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        if reduction != "mean":
            raise NotImplementedError("only reduction='mean' (the reference's, runner.py:35) is implemented")
        self.blank, self.zero_infinity = blank, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        if log_probs.dim() != 3:
            raise ValueError("expected log_probs of shape (T, B, V)")
        return ctc_loss_from_logits(log_probs.transpose(0, 1), targets, input_lengths, target_lengths, self.blank,
                                    self.zero_infinity)
