"""The callers either side of the hot path inside the reference's `ASRNN`, as
bound-method drop-ins that `patch_asrnn(model, encoder=True)` installs on a reference ASRNN instance:

    encoder(self, x, input_lens, SpecAugment=False, finetuning=False)      /root/reference/lib/standard/asrnn.py:193-221
        SpecAugment -> conv_sub_sampling -> flatten -> standard_linear -> view -> dropout -> drop zero-length rows, crop to
        the longest utterance -> conformers -> pad back -> flatten(0,1) -> projection_block
    projection_block(self, x, finetuning=False)                            asrnn.py:73-89
        Linear -> SiLU -> BatchNorm1d over the batch_size*max_len rows

Everything uses the model's OWN parameters (`standard_linear`, `projection_fc[_1]`, `projection_batch_norm`), so state_dict,
optimizer and checkpoints are untouched.  What runs where:
  * bridge: ONE tcgen05 GEMM [B, out_size] x [out_size, 512*max_len] on the token-major conv-subsampling output -- the
    flatten order of the reference (c, f', t') is absorbed into a permuted bf16 copy of the weight (refreshed when the
    parameter changes / always under CUDA-graph capture), the dropout (asrnn.py:210) is the GEMM's epilogue;
  * length filter: the reference reads `torch.max(input_lens).item()` (a device sync, asrnn.py:213).  Pass the lengths as a
    HOST list / CPU tensor (`lens_host=`, or a CPU `input_lens`) and nothing synchronises -- the whole forward is then
    capturable in a CUDA graph; with device lengths only, one `.tolist()` keeps the reference's exact semantics;
  * projection block: GEMM with SiLU epilogue + the BatchNorm1d kernels of csrc/bn1d.cu.
"""
from __future__ import annotations

import torch

from . import _lib, ops


# ------------------------------------------------------------------------------------------------------------------
# projection block
# ------------------------------------------------------------------------------------------------------------------
class _ProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn, training):
        K, C = weight.shape[1], weight.shape[0]
        x2 = x.detach().reshape(-1, K).to(torch.float32).contiguous()
        R = x2.shape[0]
        a = ops.scale_cast(x2, 1.0)
        w_bf16 = ops.cast_bf16(weight.detach().contiguous())
        z = torch.empty((R, C), device=x.device, dtype=torch.bfloat16)
        h = ops.gemm(a, w_bf16, M=R, N=C, K=K, bias=bias, act=ops.ACT_SILU, out2=z)         # asrnn.py:84-87
        if training:
            stats = torch.zeros(2 * C, device=x.device, dtype=torch.float64)
            _lib.call("b2c_bn1d_stats", h.data_ptr(), stats.data_ptr(), R, C, _lib.stream_ptr())
            mean, rstd = ops.convmod_bn_finalize(stats, R, bn.running_mean, bn.running_var, bn.eps,
                                                 bn.momentum if bn.momentum is not None else 0.1)
            bn.num_batches_tracked += 1
        else:
            mean, rstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
        y = torch.empty((R, C), device=x.device, dtype=torch.float32)
        _lib.call("b2c_bn1d_apply", h.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                  y.data_ptr(), R, C, _lib.stream_ptr())                                      # asrnn.py:88
        ctx.save_for_backward(a, w_bf16, z, h, mean, rstd, gamma)
        ctx.training, ctx.in_shape = training, x.shape
        return y.view(*x.shape[:-1], C)

    @staticmethod
    def backward(ctx, dy):
        a, w_bf16, z, h, mean, rstd, gamma = ctx.saved_tensors
        R, K = a.shape
        C = w_bf16.shape[0]
        dy2 = dy.reshape(R, C).to(torch.float32).contiguous()
        sums = torch.zeros(2 * C, device=dy.device, dtype=torch.float64)
        _lib.call("b2c_bn1d_bwd_reduce", dy2.data_ptr(), h.data_ptr(), mean.data_ptr(), rstd.data_ptr(), sums.data_ptr(), R, C,
                  _lib.stream_ptr())
        db = torch.zeros(C, device=dy.device, dtype=torch.float32)
        dz = torch.empty((R, C), device=dy.device, dtype=torch.bfloat16)
        _lib.call("b2c_bn1d_bwd_apply", dy2.data_ptr(), h.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                  gamma.data_ptr(), sums.data_ptr(), int(ctx.training), dz.data_ptr(), db.data_ptr(), R, C, _lib.stream_ptr())
        dW = torch.zeros((C, K), device=dy.device, dtype=torch.float32)
        ops.gemm(dz, a, M=C, N=K, K=R, a_mn=True, b_mn=True, out=dW, accumulate=True,
                 split_k=max(1, min(32, ((R + 63) // 64) // 4)))
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dz, w_bf16, M=R, N=K, K=C, b_mn=True, out_dtype=torch.float32).view(ctx.in_shape)
        return dx, dW, db, sums[C:].to(torch.float32), sums[:C].to(torch.float32), None, None


def projection_block(self, x, finetuning: bool = False):
    """asrnn.py:73-89 on the model's own `projection_fc` / `projection_fc_1` / `projection_batch_norm`."""
    _lib.require_device()
    fc = self.projection_fc_1 if finetuning else self.projection_fc
    bn = self.projection_batch_norm
    if not x.is_cuda:
        raise _lib.B2CError("projection_block needs CUDA tensors; there is no CPU fallback")
    return _ProjFn.apply(x, fc.weight, fc.bias, bn.weight, bn.bias, bn, bool(bn.training))


# ------------------------------------------------------------------------------------------------------------------
# bridge (flatten -> standard_linear -> view -> dropout)
# ------------------------------------------------------------------------------------------------------------------
class _FlatBridgeFn(torch.autograd.Function):
    """y[B, N] = dropout(a[B, K] Wp^T + b): a = token-major conv-subsampling output (bf16), Wp = the weight with its input
    features permuted from the reference's (c, f', t') flatten order to the kernel's (t', f', c) order."""

    @staticmethod
    def forward(ctx, a, weight, bias, w_bf16, perm, drop_p, drop_seed):
        Bn, K = a.shape
        N = weight.shape[0]
        ctx.save_for_backward(a, w_bf16)
        ctx.meta = (Bn, K, N, perm, drop_p, drop_seed)
        ctx.epoch = _lib.epoch_snapshot(a.device) if drop_p > 0.0 else None
        with _lib.epoch_scope(ctx.epoch):
            return ops.gemm(a, w_bf16, M=Bn, N=N, K=K, bias=bias, out_dtype=torch.float32, drop_p=drop_p, drop_seed=drop_seed)

    @staticmethod
    def backward(ctx, dy):
        a, w_bf16 = ctx.saved_tensors
        Bn, K, N, (C, Fo, To), drop_p, drop_seed = ctx.meta
        db = torch.zeros(N, device=dy.device, dtype=torch.float32)
        # the dropout index is the LINEAR element index b*N + n, so the mask can be regenerated on any row-major reshape
        # of dy: N = 512 * max_len exceeds the cast kernel's column limit, [B * k, N / k] does not
        k = 1
        while N // k > 8192 or N % k:
            k += 1
        with _lib.epoch_scope(ctx.epoch):
            dyb = ops.scale_cast(dy.contiguous().view(Bn * k, N // k), 1.0, drop_p, drop_seed).view(Bn, N)
        for c0 in range(0, N, 8192):                                            # bias gradient: column sums over the batch
            c1 = min(N, c0 + 8192)
            ops.colsum(dyb[:, c0:c1], db[c0:c1])
        dWp = torch.zeros((N, K), device=dy.device, dtype=torch.float32)
        ops.gemm(dyb, a, M=N, N=K, K=Bn, a_mn=True, b_mn=True, out=dWp, accumulate=True)
        dW = dWp.view(N, To, Fo, C).permute(0, 3, 2, 1).reshape(N, K)          # back to the reference's (c, f', t') feature order
        da = ops.gemm(dyb, w_bf16, M=Bn, N=K, K=N, b_mn=True) if ctx.needs_input_grad[0] else None
        return da, dW, db, None, None, None, None


def _bridge_weight(self, C, Fo, To):
    """bf16 copy of `standard_linear.weight` with permuted input features, in a static buffer (same refresh policy as
    conformer._WeightCache: on parameter change, and always while the stream is being captured)."""
    w = self.standard_linear.weight
    st = self.__dict__.setdefault("_b2c_bridge", {})
    key = (w.data_ptr(), w._version)
    if st.get("buf") is None or st["buf"].device != w.device or st.get("ptr") != key[0]:
        st["buf"] = torch.empty(w.shape, device=w.device, dtype=torch.bfloat16)
        st["ptr"], st["key"] = key[0], None
    if key != st["key"] or torch.cuda.is_current_stream_capturing():
        with torch.no_grad():            # permute + cast in one strided copy (weight preparation, not data path)
            st["buf"].view(-1, To, Fo, C).copy_(w.detach().view(-1, C, Fo, To).permute(0, 3, 2, 1))
        st["key"] = key
    return st["buf"]


def encoder(self, x, input_lens, SpecAugment: bool = False, finetuning: bool = False, lens_host=None):
    """Drop-in for `ASRNN.encoder` (asrnn.py:193-221).  x (B,1,F,T) fp32 CUDA; input_lens (B,) (zero for batch-padding rows,
    which the reference's loader appends at the END of a batch, speechcommands.py:189-190)."""
    _lib.require_device()
    hp = self.hp
    if SpecAugment:
        x = self.SpecAugment(x, lens_host if lens_host is not None else input_lens)         # asrnn.py:204-205
    cs = self.conv_sub_sampling
    tok = cs.forward_tokens(x)                                                              # (B, T', F', C) bf16
    B, To, Fo, C = tok.shape
    w_bf16 = _bridge_weight(self, C, Fo, To)
    N = self.standard_linear.weight.shape[0]
    drop_p = float(self.dropout.p) if self.training else 0.0
    drop_seed = (self.conformers.next_seed() ^ 0x5DEECE66D) if drop_p > 0.0 else 0
    h = _FlatBridgeFn.apply(tok.reshape(B, To * Fo * C), self.standard_linear.weight, self.standard_linear.bias, w_bf16,
                            (C, Fo, To), drop_p, drop_seed)                                 # asrnn.py:207-210
    h = h.view(hp.batch_size, hp.max_len, N // hp.max_len)
    # asrnn.py:211-213: drop the zero-length rows (a suffix of the batch), crop to the longest utterance
    if lens_host is None:
        lens_host = input_lens.tolist() if isinstance(input_lens, torch.Tensor) else list(input_lens)
    lens_host = [int(v) for v in lens_host]
    n = sum(1 for v in lens_host if v > 0)
    if any(v > 0 for v in lens_host[n:]):
        raise ValueError("zero-length rows must follow the real utterances (the reference pairs x[:n] with the compacted "
                         "lengths, asrnn.py:211-213)")
    t_max = max(lens_host[:n]) if n else 0
    lens_dev = input_lens[:n] if isinstance(input_lens, torch.Tensor) and input_lens.is_cuda else \
        torch.tensor(lens_host[:n], device=x.device)
    y, out_lens = self.conformers(h[:n, :t_max, :].contiguous(), lens_dev)                  # asrnn.py:214
    y = torch.nn.functional.pad(y, (0, 0, 0, hp.max_len - y.shape[1], 0, hp.batch_size - n))    # asrnn.py:215
    y = y.flatten(0, 1)
    y = self.projection_block(y)                                                            # asrnn.py:217
    if finetuning and hp.extra_proj:
        y = self.dropout(y)
        y = self.projection_block(y, True)
    return y, out_lens
