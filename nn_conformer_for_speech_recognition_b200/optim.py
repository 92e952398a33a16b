"""Fused Adafactor: the reference's optimizer (`transformers.Adafactor(lr, beta1=0.9, scale_parameter=False,
relative_step=False)`, /root/reference/lib/standard/runner.py:36) as four multi-tensor kernel launches per step
(csrc/optim.cu) instead of ~15 small launches per parameter.  Same update rule and state (exp_avg, factored
exp_avg_sq_row / exp_avg_sq_col for >= 2-D tensors, exp_avg_sq for vectors), eps=(1e-30, 1e-3), clip_threshold=1.0,
decay_rate=-0.8, no weight decay -- HF's defaults, which the reference does not override.
"""
from __future__ import annotations

from typing import Iterable

import numpy as np
import torch

from . import _lib

_TENSOR_DT = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("row", "<u8"), ("col", "<u8"), ("rowsum", "<u8"),
                       ("colsum", "<u8"), ("rmean", "<u8"), ("sumsq", "<u8"), ("P", "<i4"), ("R", "<i4"), ("C", "<i4"),
                       ("factored", "<i4"), ("n", "<i8")])
_CHUNK_DT = np.dtype([("t", "<i4"), ("row0", "<i4"), ("nrows", "<i4"), ("pad", "<i4")])
assert _TENSOR_DT.itemsize == 96 and _CHUNK_DT.itemsize == 16
_CHUNK_ELEMS = 16384
_CHUNK_ELEMS_NARROW = 2048      # rows shorter than a warp run one thread per row: keep the per-thread loops short


class Adafactor:
    """Minimal optimizer interface (step / zero_grad / state) -- enough for `Runner.train`'s loop (runner.py:140-146)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2e-5, beta1: float = 0.9,
                 eps=(1e-30, 1e-3), clip_threshold: float = 1.0, decay_rate: float = -0.8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("fused Adafactor needs contiguous fp32 CUDA parameters (no fallback)")
        self.lr, self.beta1, self.eps1, self.clip, self.decay_rate = lr, beta1, eps[0], clip_threshold, decay_rate
        self.step_count = 0                               # host mirror; the authoritative counter lives on the device
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self._beta2t_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        geo = []
        for p in self.params:
            if p.dim() >= 2:
                C, R = p.shape[-1], p.shape[-2]
                geo.append((p.numel() // (R * C), R, C, 1))
            else:
                geo.append((1, 1, p.numel(), 0))
        al = lambda n: (n + 3) // 4 * 4                                             # noqa: E731
        n_m = sum(al(p.numel()) for p in self.params)
        n_row = sum(al(P * R) if f else al(P * R * C) for P, R, C, f in geo)
        n_col = sum(al(P * C) for P, R, C, f in geo if f)
        n_ws = sum(al(P * R) + al(P * C) + al(P) + 4 for P, R, C, f in geo)
        self.m = torch.zeros(n_m, device=dev)
        self.row = torch.zeros(n_row, device=dev)
        self.col = torch.zeros(max(n_col, 4), device=dev)
        self.ws = torch.zeros(n_ws, device=dev)          # rowsum | colsum | rmean | sumsq, zeroed every step
        tab = np.zeros(len(self.params), _TENSOR_DT)
        chunks = []
        om = orow = ocol = ows = 0
        self.views = []
        for i, (p, (P, R, C, f)) in enumerate(zip(self.params, geo)):
            t = tab[i]
            t["p"] = p.data_ptr()
            t["m"] = self.m.data_ptr() + 4 * om
            t["row"] = self.row.data_ptr() + 4 * orow
            t["col"] = self.col.data_ptr() + 4 * ocol
            t["rowsum"] = self.ws.data_ptr() + 4 * ows
            t["colsum"] = t["rowsum"] + 4 * al(P * R)
            t["rmean"] = t["colsum"] + 4 * al(P * C)
            t["sumsq"] = t["rmean"] + 4 * al(P)
            t["P"], t["R"], t["C"], t["factored"], t["n"] = P, R, C, f, p.numel()
            self.views.append(dict(exp_avg=self.m[om:om + p.numel()].view(p.shape),
                                   row=self.row[orow:orow + (P * R if f else p.numel())],
                                   col=self.col[ocol:ocol + P * C] if f else None))
            om += al(p.numel())
            orow += al(P * R) if f else al(P * R * C)
            ocol += al(P * C) if f else 0
            ows += al(P * R) + al(P * C) + al(P) + 4
            rows = P * R
            per = max(1, (_CHUNK_ELEMS if C >= 32 else _CHUNK_ELEMS_NARROW) // C)
            for r0 in range(0, rows, per):
                chunks.append((i, r0, min(per, rows - r0), 0))
        self._tab_host = torch.from_numpy(tab.view(np.uint8).reshape(-1)).pin_memory()
        self._tab_np = self._tab_host.numpy().view(_TENSOR_DT)
        self._tab_dev = torch.empty_like(self._tab_host, device=dev)
        ck = np.array(chunks, dtype=np.int32).reshape(-1, 4)
        self._chunks_dev = torch.from_numpy(ck.view(np.uint8).reshape(-1)).to(dev)
        self.n_chunks = len(chunks)
        self._ptrs = None

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        grads = []
        for p in self.params:
            g = p.grad
            if g is None:
                raise RuntimeError("fused Adafactor: every parameter needs a gradient (the reference's loop always has one)")
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.to(torch.float32).contiguous()
                p.grad = g
            grads.append(g)
        ptrs = [g.data_ptr() for g in grads]
        if ptrs != self._ptrs:                             # gradient buffers moved: refresh the device table
            self._tab_np["g"] = ptrs
            self._tab_dev.copy_(self._tab_host, non_blocking=True)
            self._ptrs = ptrs
        self.step_count += 1
        self.ws.zero_()
        _lib.call("b2c_adafactor_step", self._tab_dev.data_ptr(), len(self.params), self._chunks_dev.data_ptr(),
                  self.n_chunks, self._step_dev.data_ptr(), self._beta2t_dev.data_ptr(), self.lr, self.beta1, self.eps1,
                  self.clip, self.decay_rate, _lib.stream_ptr())
        # the kernels wrote the parameters through raw pointers: tell autograd / the bf16 weight caches they changed
        torch.autograd.graph.increment_version(self.params)
