"""Drop-in for `ASRNN.SpecAugment / time_warping / frequency_masking / time_masking`
(/root/reference/lib/standard/asrnn.py:91-192).

The random draws are made on the host with the interpreter's GLOBAL `random` module, in exactly the order the
reference makes them -- so `random.seed(s)` before the call gives the same parameters and
leaves the stream in the same state as the reference would.  The data path is one sm_100a kernel
(csrc/specaug.cu): warp gather + masks, no host round trip of the spectrogram.

mode="reference" (default): bit-exact with the code as it runs -- the two mask assignments at asrnn.py:141,165 hit
    temporaries, so the frequency/time masks are drawn but never applied.
mode="paper": same draws, masks applied as the docstrings / main.ipynb cell 24 describe.
Functions take `self` (anything with an `.hp`) so they can be bound onto an ASRNN instance (see patch.py).
"""
from __future__ import annotations

import random
from math import floor

import torch

from . import _lib, ops

MODE = "reference"      # module-level default; patch_asrnn(model, specaug_mode=...) overrides per model


def _mode(self):
    return getattr(self, "_b2c_specaug_mode", MODE)


def _tau_list(tau):
    if isinstance(tau, torch.Tensor):
        return [int(v) for v in tau.tolist()]
    return [int(v) for v in tau]


def _i32(values, device):
    return torch.tensor(values, dtype=torch.int32).to(device, non_blocking=True)


def _draw_warp(hp, B, tau):
    W = hp.warping_param_W
    ws, w0s = [], []
    for u in range(B):
        w = random.randint(-W, W)                                             # asrnn.py:103
        w0 = W if tau[u] < 2 * W else random.randint(W, tau[u] - W - 1)       # asrnn.py:104-107 (ValueError if tau==2W)
        if tau[u] > 0 and w0 == 0:
            raise ZeroDivisionError("division by zero")                       # (w0+w)/w0 at asrnn.py:110
        ws.append(w)
        w0s.append(w0)
    return ws, w0s


def _draw_fmask(hp, v):
    f = random.randint(0, hp.frequency_mask_param_F)                          # asrnn.py:138
    f0 = random.randint(0, v - hp.frequency_mask_param_F)                     # asrnn.py:139
    return f0, f


def _draw_tmask(hp, B, tau):
    out = []
    for u in range(B):
        T_ = floor(hp.ps * tau[u]) if (u < len(tau) and hp.adaptive_size) else hp.time_mask_param_T
        t = random.randint(0, T_)                                             # asrnn.py:162
        t0 = random.randint(0, max(tau[u] - T_, tau[u]))                      # asrnn.py:163
        out.append((t0, t))
    return out


def _launch(x3, tau, warps, fmasks, tmasks, apply_masks, mask_value):
    dev = x3.device
    B = x3.shape[0]
    tau_t = _i32(tau + [0] * (B - len(tau)), dev)
    ww = _i32([w[0] for w in warps], dev) if warps else None
    w0 = _i32([w[1] for w in warps], dev) if warps else None
    fm = _i32(fmasks, dev) if fmasks else None
    tm = _i32(tmasks, dev) if tmasks else None
    return ops.specaug_apply(x3.to(torch.float32).contiguous(), tau_t, ww, w0, fm, tm, apply_masks, float(mask_value))


def time_warping(self, x, tau):
    """asrnn.py:91-125: x (B,F,T) -> (B,F,T) fp32 on the input's device."""
    tau = _tau_list(tau)
    return _launch(x, tau, [_draw_warp(self.hp, x.shape[0], tau)], [], [], False, 0.0)


def frequency_masking(self, x):
    """asrnn.py:127-144: draws (f, f0); masks rows only in "paper" mode."""
    f0, f = _draw_fmask(self.hp, x.shape[1])
    if _mode(self) == "reference":
        return x.clone()
    return _launch(x, [x.shape[2]] * x.shape[0], [], [(f0, f)], [], True, self.hp.mask_value)


def time_masking(self, x, tau):
    """asrnn.py:146-168."""
    tau = _tau_list(tau)
    tm = _draw_tmask(self.hp, x.shape[0], tau)
    if _mode(self) == "reference":
        return x.clone()
    return _launch(x, tau, [], [], [tm], True, self.hp.mask_value)


def SpecAugment(self, x, tau):
    """asrnn.py:170-192: (B,1,F,T) -> (B,1,F,T); all passes fused into one kernel launch."""
    _lib.require_device()
    hp = self.hp
    x = x.squeeze()
    if len(list(x.shape)) < 3:
        x = x.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError(f"SpecAugment expects (B,1,F,T) with F,T > 1; got squeezed shape {tuple(x.shape)}")
    B, F, _ = x.shape
    tau = _tau_list(tau)
    warps = [_draw_warp(hp, B, tau) for _ in range(hp.warping_ntimes)]                     # asrnn.py:183-184
    fmasks = [_draw_fmask(hp, F) for _ in range(hp.frequency_mask_ntimes)]                 # asrnn.py:185-186
    Mt = hp.time_multiplicity
    if hp.adaptive_multiplicity:
        Mt = min(Mt, floor(hp.pm))                                                          # asrnn.py:188-189
    tmasks = [_draw_tmask(hp, B, tau) for _ in range(Mt)]                                   # asrnn.py:190-191
    out = _launch(x, tau, warps, fmasks, tmasks, _mode(self) == "paper", hp.mask_value)
    return out.unsqueeze(1)


class SpecAugPlan:
    """Pre-drawn SpecAugment for a CUDA-graph-captured step.

    `draw(...)` runs on the host BEFORE the replay (same draws, same order, same global `random` stream as
    `SpecAugment`), writes the parameters into a pinned block and enqueues its host->device copy on the current
    stream; `apply(x)` is only the kernel launch, so it can be captured once and replayed with new parameters every
    step.  The pinned block is DOUBLE-BUFFERED: in a device-resident loop nothing synchronises host and device, so the
    host can be drawing step n+1 while step n's copy has not executed yet -- `draw` alternates two pinned slots and
    waits for the copy that last read the slot it is about to overwrite (an event recorded behind that copy).  The
    device block needs no second copy: the stream orders copy(n+1) behind the kernel of step n.
    """

    def __init__(self, hp, B: int, device, mode: str = "reference"):
        self.hp, self.B, self.mode = hp, B, mode
        Mt = hp.time_multiplicity
        if hp.adaptive_multiplicity:
            Mt = min(Mt, floor(hp.pm))
        self.n_warp, self.n_f, self.n_t = hp.warping_ntimes, hp.frequency_mask_ntimes, Mt
        n = B + 2 * self.n_warp * B + 2 * self.n_f + 2 * self.n_t * B
        self.hosts = [torch.zeros(max(n, 1), dtype=torch.int32).pin_memory() for _ in range(2)]
        self.copied = [None, None]                       # event behind the last copy that read hosts[i]
        self.slot = 0
        self.dev = torch.zeros(max(n, 1), dtype=torch.int32, device=device)
        o = 0
        self.sl = {}
        for name, cnt in (("tau", B), ("w", self.n_warp * B), ("w0", self.n_warp * B), ("fm", 2 * self.n_f),
                          ("tm", 2 * self.n_t * B)):
            self.sl[name] = (o, o + cnt)
            o += cnt

    @property
    def host(self):
        """the pinned slot written by the latest draw()"""
        return self.hosts[self.slot ^ 1]

    def draw(self, F: int, tau):
        hp, B = self.hp, self.B
        tau = _tau_list(tau)
        warps = [_draw_warp(hp, B, tau) for _ in range(self.n_warp)]
        fmasks = [_draw_fmask(hp, F) for _ in range(self.n_f)]
        tmasks = [_draw_tmask(hp, B, tau) for _ in range(self.n_t)]
        i = self.slot
        if self.copied[i] is not None:
            self.copied[i].synchronize()                 # the copy issued two draws ago has read this slot
        h = self.hosts[i]

        def put(name, values):
            a, b = self.sl[name]
            if b > a:
                h[a:b] = torch.tensor(values, dtype=torch.int32).reshape(-1)
        put("tau", tau)
        put("w", [w[0] for w in warps])
        put("w0", [w[1] for w in warps])
        put("fm", fmasks)
        put("tm", tmasks)
        self.dev.copy_(h, non_blocking=True)             # ordered behind the previous step's kernel on this stream
        ev = torch.cuda.Event()
        ev.record()
        self.copied[i] = ev
        self.slot = i ^ 1

    def apply(self, x4: torch.Tensor) -> torch.Tensor:
        """x4 (B,1,F,T) fp32 on the GPU -> (B,1,F,T); one kernel launch, capturable."""
        x = x4[:, 0] if x4.dim() == 4 else x4
        d = self.dev

        def view(name, shape):
            a, b = self.sl[name]
            return d[a:b].view(shape) if b > a else None
        out = ops.specaug_apply(x.to(torch.float32).contiguous(), view("tau", (self.B,)), view("w", (self.n_warp, self.B)),
                                view("w0", (self.n_warp, self.B)), view("fm", (self.n_f, 2)), view("tm", (self.n_t, self.B, 2)),
                                self.mode == "paper", float(self.hp.mask_value))
        return out.unsqueeze(1)
