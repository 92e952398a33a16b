"""The reference's input pipeline on the GPU -- csrc/frontend.cu.

* `LogMelFrontend`: `librosa.feature.melspectrogram(y, sr, n_mels)` -> `log` -> min-max normalisation, the three steps of
  `SpeechCommands.get_mels` (/root/reference/lib/standard/speechcommands.py:113-119), for a whole batch of raw waveforms in
  two kernel launches, written straight into the zero-padded `(batch_size, n_mels, max_len)` tensor the model consumes
  (the padding of `get_batch`, speechcommands.py:186-190).  librosa itself is not installed in this image, so parity is
  pinned to its published algorithm through torchaudio's `MelSpectrogram(norm="slaney", mel_scale="slaney")` (documented
  to reproduce librosa) in tests/test_frontend_gpu.py -- "unpinned against librosa itself" is stated there.
* `PaddedBatcher`: `SpeechCommands.get_batch`'s python-list padding (speechcommands.py:174-197) for pre-computed features:
  ragged utterances are packed back to back in ONE pinned staging buffer, cross PCIe in one copy, and a kernel scatters
  them into the zero-padded batch (rows beyond the real utterances are zero, `tau` is zero there).
No CPU fallback for the data path; the mel filterbank (a constant) is built once on the host.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import _lib


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax: float = None) -> np.ndarray:
    """librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels, htk=False, norm='slaney') restated: (n_mels, n_fft//2+1) fp32."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).astype(np.float32)


class LogMelFrontend:
    def __init__(self, sr: int = 16000, n_mels: int = 40, n_fft: int = 2048, hop_length: int = 512, device="cuda"):
        if torch.device(device).type != "cuda":
            raise ValueError("LogMelFrontend needs a CUDA device (there is no CPU fallback)")
        self.sr, self.n_mels, self.n_fft, self.hop, self.device = sr, n_mels, n_fft, hop_length, torch.device(device)
        fb = mel_filterbank(sr, n_fft, n_mels)
        nz = fb > 0
        lo = np.where(nz.any(1), nz.argmax(1), 0).astype(np.int32)
        hi = np.where(nz.any(1), fb.shape[1] - nz[:, ::-1].argmax(1), 0).astype(np.int32)
        self.basis = torch.from_numpy(fb).to(self.device)
        self.f_lo, self.f_hi = torch.from_numpy(lo).to(self.device), torch.from_numpy(hi).to(self.device)

    def frames(self, n_samples: int) -> int:
        return n_samples // self.hop + 1

    def __call__(self, waves: Sequence, batch_size: int = None, max_len: int = None):
        """waves: list of 1-D float waveforms (numpy / torch, host or device).  Returns (mels (B, n_mels, max_len) fp32 on the
        GPU, frames (B,) int32: frames per utterance, 0 for the batch-padding rows)."""
        _lib.require_device()
        n = len(waves)
        B = n if batch_size is None else batch_size
        if B < n or B < 1:
            raise ValueError("batch_size must be >= the number of waveforms")
        lens = [int(w.shape[-1]) for w in waves]
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum(lens)
        host = torch.empty(max(int(off[-1]), 1), dtype=torch.float32).pin_memory()
        for i, w in enumerate(waves):
            wt = torch.as_tensor(w, dtype=torch.float32).reshape(-1)
            host[off[i]:off[i + 1]].copy_(wt)                                   # host-side memcpy per utterance, no Python loops over samples
        wave = host.to(self.device, non_blocking=True)
        off_d = torch.from_numpy(off).to(self.device, non_blocking=True)
        fmax = max([self.frames(v) for v in lens], default=1)
        t_max = fmax if max_len is None else max_len
        scratch = torch.empty((max(n, 1), self.n_mels, fmax), device=self.device, dtype=torch.float32)
        out = torch.empty((B, self.n_mels, t_max), device=self.device, dtype=torch.float32)
        frames = torch.empty(B, device=self.device, dtype=torch.int32)
        _lib.call("b2c_logmel", wave.data_ptr(), off_d.data_ptr(), n, B, self.basis.data_ptr(), self.f_lo.data_ptr(),
                  self.f_hi.data_ptr(), self.n_fft, self.hop, self.n_mels, fmax, t_max, scratch.data_ptr(), out.data_ptr(),
                  frames.data_ptr(), _lib.stream_ptr())
        return out, frames


class PaddedBatcher:
    """`get_batch` (speechcommands.py:174-197) for pre-computed features: list of (F, len_i) arrays -> (mels (B, F, max_len),
    tau (B,) int64) on the GPU, zero padded, in one pinned staging copy + one kernel."""

    def __init__(self, batch_size: int, n_features: int, max_len: int, device="cuda"):
        if torch.device(device).type != "cuda":
            raise ValueError("PaddedBatcher needs a CUDA device (there is no CPU fallback)")
        self.B, self.F, self.T, self.device = batch_size, n_features, max_len, torch.device(device)
        self.host = torch.empty(batch_size * n_features * max_len, dtype=torch.float32).pin_memory()

    def __call__(self, feats: Sequence, tau: Sequence[int] = None):
        _lib.require_device()
        n = len(feats)
        if n > self.B:
            raise ValueError("more utterances than batch_size")
        lens = [int(f.shape[-1]) for f in feats]
        if any(f.shape[0] != self.F for f in feats) or any(v > self.T for v in lens):
            raise ValueError(f"features must be ({self.F}, len <= {self.T})")
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum([self.F * v for v in lens])
        for i, f in enumerate(feats):
            self.host[off[i]:off[i + 1]].copy_(torch.as_tensor(f, dtype=torch.float32).reshape(-1))
        packed = self.host[:max(int(off[-1]), 1)].to(self.device, non_blocking=True)
        off_d = torch.from_numpy(off).to(self.device, non_blocking=True)
        lens_d = torch.tensor(lens + [0] * (self.B - n), dtype=torch.int32).to(self.device, non_blocking=True)
        out = torch.empty((self.B, self.F, self.T), device=self.device, dtype=torch.float32)
        _lib.call("b2c_pad_batch", packed.data_ptr(), off_d.data_ptr(), lens_d.data_ptr(), n, self.B, self.F, self.T,
                  out.data_ptr(), _lib.stream_ptr())
        # the reference's tau is len(mel) = the number of MEL ROWS (speechcommands.py:71); callers that want
        # that quirk pass it explicitly -- the default is the true frame count
        tau_t = torch.tensor(list(tau) + [0] * (self.B - n) if tau is not None else lens + [0] * (self.B - n), dtype=torch.int64)
        return out, tau_t.to(self.device, non_blocking=True)
