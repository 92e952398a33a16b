"""`LSTM`: drop-in for the reference's bi-LSTM "decoder"
    self.lstm = nn.LSTM(hp.projection_out_size, hp.standard_decoder_nodes, bidirectional=..., num_layers=1)   asrnn.py:38
    x = self.lstm(x)[0]                                                                                    asrnn.py:252
(SURVEY.md section 8f row 4).  It IS a `torch.nn.LSTM` subclass -- same constructor, parameters (`weight_ih_l0`,
`weight_hh_l0`, `bias_*_l0`, `*_reverse`), initialisation and state_dict -- whose forward / backward run on this library:

  * all input projections of a direction in ONE tcgen05 GEMM  Gx = X W_ih^T + (b_ih + b_hh)  [L*N, 4H] fp32;
  * per time step one GEMM  gates = Gx[t] + h_{t-1} W_hh^T  (fp32-residual epilogue) and one `b2c_lstm_cell_fwd`;
  * backward: per step `b2c_lstm_cell_bwd` (writes the pre-activation gradients as the bf16 rows of dG) and the recurrent
    dgrad GEMM  dh = dG[t] W_hh;  after the loop the weight gradients are three GEMMs over all L*N rows at once
    (dW_hh = dG^T H_prev, dW_ih = dG^T X, dX = dG W_ih) plus a column sum for the biases.

Like `nn.LSTM` with `batch_first=False` the first axis is time -- the reference feeds it `[batch, max_len, C]`, so its "time"
is the batch axis (asrnn.py:252); this module reproduces that, whatever the caller meant.  Operands of the GEMMs are bf16
(fp32 accumulation, fp32 cell state and gates): parity with `torch.nn.LSTM` is checked at bf16 tolerance in
tests/test_lstm_gpu.py.  Only what the reference uses is supported: one layer, no projection; anything else raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops
from .autograd import _split_k


def _cell_fwd(gates, c_prev, c_out, h_out, h_bf16, acts, N, H):
    _lib.call("b2c_lstm_cell_fwd", gates.data_ptr(), _lib.ptr(c_prev), c_out.data_ptr(), h_out.data_ptr(), h_out.stride(0),
              h_bf16.data_ptr(), _lib.ptr(acts), N, H, _lib.stream_ptr())


def _cell_bwd(dout, dh_rec, dc, acts, c_prev, c_cur, dgates, N, H):
    _lib.call("b2c_lstm_cell_bwd", _lib.ptr(dout), dout.stride(0) if dout is not None else 0, _lib.ptr(dh_rec), dc.data_ptr(),
              acts.data_ptr(), _lib.ptr(c_prev), c_cur.data_ptr(), dgates.data_ptr(), N, H, _lib.stream_ptr())


class _LSTMFn(torch.autograd.Function):
    """inputs: x [L, N, In] fp32, h0 / c0 [dirs, N, H] or None, then per direction (w_ih, w_hh, b_ih, b_hh) (biases may be None)."""

    @staticmethod
    def forward(ctx, x, h0, c0, H, dirs, save, *weights):
        L, N, In = x.shape
        dev = x.device
        x2 = x.detach().reshape(L * N, In).to(torch.float32).contiguous()
        xb = ops.scale_cast(x2, 1.0)
        out = torch.empty((L, N, dirs * H), device=dev, dtype=torch.float32)
        h_n = torch.empty((dirs, N, H), device=dev, dtype=torch.float32)
        c_n = torch.empty((dirs, N, H), device=dev, dtype=torch.float32)
        gates_buf = torch.empty((N, 4 * H), device=dev, dtype=torch.float32)
        saved = []
        for d in range(dirs):
            w_ih, w_hh, b_ih, b_hh = weights[4 * d:4 * d + 4]
            w_ih_b = ops.cast_bf16(w_ih.detach().contiguous())
            w_hh_b = ops.cast_bf16(w_hh.detach().contiguous())
            bsum = None
            if b_ih is not None:
                bsum = (b_ih.detach() + b_hh.detach()).to(torch.float32).contiguous()
            gx = ops.gemm(xb, w_ih_b, M=L * N, N=4 * H, K=In, bias=bsum, out_dtype=torch.float32).view(L, N, 4 * H)
            c_all = torch.empty((L, N, H), device=dev, dtype=torch.float32)
            hb_all = torch.empty((L, N, H), device=dev, dtype=torch.bfloat16)
            acts = torch.empty((L, N, 4 * H), device=dev, dtype=torch.float32) if save else None
            h0b = ops.scale_cast(h0[d].detach().to(torch.float32).contiguous(), 1.0) if h0 is not None else None
            c0d = c0[d].detach().to(torch.float32).contiguous() if c0 is not None else None
            order = range(L) if d == 0 else range(L - 1, -1, -1)
            tp = None
            for t in order:
                hprev = hb_all[tp] if tp is not None else h0b
                if hprev is None:
                    gates = gx[t]
                else:
                    gates = ops.gemm(hprev, w_hh_b, M=N, N=4 * H, K=H, resid=gx[t], out=gates_buf)
                _cell_fwd(gates, c_all[tp] if tp is not None else c0d, c_all[t], out[t, :, d * H:(d + 1) * H], hb_all[t],
                          acts[t] if save else None, N, H)
                tp = t
            h_n[d].copy_(out[tp, :, d * H:(d + 1) * H])
            c_n[d].copy_(c_all[tp])
            saved.append((w_ih_b, w_hh_b, c_all, hb_all, acts, h0b, c0d))
        ctx.saved = saved if save else None
        ctx.xb, ctx.dims, ctx.has_bias = xb, (L, N, In, H, dirs), weights[2] is not None
        ctx.has_h0 = h0 is not None
        ctx.x_dtype = x.dtype
        return out, h_n, c_n

    @staticmethod
    def backward(ctx, dout, dh_n, dc_n):
        L, N, In, H, dirs = ctx.dims
        xb = ctx.xb
        dev = xb.device
        dout = dout.to(torch.float32).contiguous() if dout is not None else None
        dx = None
        wgrads = []
        dh0 = torch.zeros((dirs, N, H), device=dev, dtype=torch.float32) if ctx.has_h0 else None
        dc0 = torch.zeros((dirs, N, H), device=dev, dtype=torch.float32) if ctx.has_h0 else None
        for d in range(dirs):
            w_ih_b, w_hh_b, c_all, hb_all, acts, h0b, c0d = ctx.saved[d]
            dg_all = torch.empty((L, N, 4 * H), device=dev, dtype=torch.bfloat16)
            dc = dc_n[d].to(torch.float32).clone().contiguous() if dc_n is not None else torch.zeros((N, H), device=dev, dtype=torch.float32)
            dh_rec = dh_n[d].to(torch.float32).contiguous() if dh_n is not None else None
            dh_buf = torch.empty((N, H), device=dev, dtype=torch.float32)
            order = list(range(L)) if d == 0 else list(range(L - 1, -1, -1))
            for s in range(L - 1, -1, -1):
                t = order[s]
                tp = order[s - 1] if s > 0 else None
                _cell_bwd(dout[t, :, d * H:(d + 1) * H] if dout is not None else None, dh_rec, dc, acts[t],
                          c_all[tp] if tp is not None else c0d, c_all[t], dg_all[t], N, H)
                if s > 0 or ctx.has_h0:
                    dh_rec = ops.gemm(dg_all[t], w_hh_b, M=N, N=H, K=4 * H, b_mn=True, out=dh_buf)
            if ctx.has_h0:
                dh0[d].copy_(dh_rec)
                dc0[d].copy_(dc)
            dg2 = dg_all.view(L * N, 4 * H)
            first = h0b if h0b is not None else torch.zeros((N, H), device=dev, dtype=torch.bfloat16)
            if d == 0:
                hprev = torch.cat([first.unsqueeze(0), hb_all[:-1]], 0)
            else:
                hprev = torch.cat([hb_all[1:], first.unsqueeze(0)], 0)
            hprev2 = hprev.view(L * N, H)
            tok = L * N
            dw_hh = ops.gemm(dg2, hprev2, M=4 * H, N=H, K=tok, a_mn=True, b_mn=True, out_dtype=torch.float32, accumulate=True,
                             split_k=_split_k(tok, ((4 * H + 127) // 128) * ((H + 255) // 256)))
            dw_ih = ops.gemm(dg2, xb, M=4 * H, N=In, K=tok, a_mn=True, b_mn=True, out_dtype=torch.float32, accumulate=True,
                             split_k=_split_k(tok, ((4 * H + 127) // 128) * ((In + 255) // 256)))
            db = None
            if ctx.has_bias:
                db = ops.colsum(dg2, torch.zeros(4 * H, device=dev, dtype=torch.float32))
            dxd = ops.gemm(dg2, w_ih_b, M=tok, N=In, K=4 * H, b_mn=True, out_dtype=torch.float32)
            dx = dxd if dx is None else dx + dxd
            wgrads += [dw_ih, dw_hh, db, db]
        ctx.saved = None
        return (dx.view(L, N, In).to(ctx.x_dtype), dh0, dc0, None, None, None, *wgrads)


class LSTM(torch.nn.LSTM):
    """torch.nn.LSTM (one layer, optional bidirectional, no projection) computing on libb200conformer.so."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.num_layers != 1 or self.proj_size != 0:
            raise NotImplementedError("b2c.LSTM covers what the reference builds: num_layers=1, proj_size=0 (asrnn.py:38, hparams.py:77)")
        if self.input_size % 8 or self.hidden_size % 8:
            raise NotImplementedError("b2c.LSTM: input_size and hidden_size must be multiples of 8 (16-byte GEMM rows)")

    def flatten_parameters(self) -> None:          # cuDNN's packed weight buffer is not used
        return

    def forward(self, input, hx: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):  # noqa: A002
        _lib.require_device()
        if isinstance(input, torch.nn.utils.rnn.PackedSequence):
            raise NotImplementedError("b2c.LSTM: packed sequences are not used by the reference (asrnn.py:252)")
        if input.dim() != 3:
            raise ValueError("b2c.LSTM expects a 3-D input [L, N, input_size] (or [N, L, input_size] with batch_first)")
        x = input.transpose(0, 1) if self.batch_first else input
        if x.shape[2] != self.input_size:
            raise ValueError(f"b2c.LSTM: input feature size {x.shape[2]} != input_size {self.input_size}")
        dirs = 2 if self.bidirectional else 1
        h0 = c0 = None
        if hx is not None:
            h0, c0 = hx
            want = (dirs, x.shape[1], self.hidden_size)
            if tuple(h0.shape) != want or tuple(c0.shape) != want:
                raise ValueError(f"b2c.LSTM: hx tensors must have shape {want}")
        weights = []
        for sfx in (("",) if dirs == 1 else ("", "_reverse")):
            weights += [getattr(self, f"weight_ih_l0{sfx}"), getattr(self, f"weight_hh_l0{sfx}"),
                        getattr(self, f"bias_ih_l0{sfx}") if self.bias else None,
                        getattr(self, f"bias_hh_l0{sfx}") if self.bias else None]
        save = torch.is_grad_enabled() and (x.requires_grad or any(w is not None and w.requires_grad for w in weights) or
                                            (h0 is not None and (h0.requires_grad or c0.requires_grad)))
        out, h_n, c_n = _LSTMFn.apply(x.contiguous(), h0, c0, self.hidden_size, dirs, save, *weights)
        if self.batch_first:
            out = out.transpose(0, 1)
        return out.to(input.dtype), (h_n.to(input.dtype), c_n.to(input.dtype))


def lstm_from_reference(ref: torch.nn.LSTM) -> LSTM:
    new = LSTM(ref.input_size, ref.hidden_size, num_layers=ref.num_layers, bias=ref.bias, batch_first=ref.batch_first,
               dropout=ref.dropout, bidirectional=ref.bidirectional, proj_size=ref.proj_size)
    new.load_state_dict(ref.state_dict())
    p = next(ref.parameters())
    return new.to(device=p.device, dtype=p.dtype).train(ref.training)
