"""Input pipeline on the GPU -- csrc/frontend.cu.

Oracle for the log-mel front end: librosa is NOT installed in this image, so parity against librosa
itself is UNPINNED; the published algorithm is pinned through torchaudio's `MelSpectrogram(n_fft=2048, hop_length=512,
power=2, center=True, pad_mode="constant", norm="slaney", mel_scale="slaney")` (documented to reproduce
`librosa.feature.melspectrogram`) in fp64, followed by the reference's own three lines (speechcommands.py:114,117-119).
Tolerance: |d| <= 2e-4 on the min-max-normalised [0,1] output (fp32 FFT with fp32 twiddles; the log compresses errors of
the small bins, the normalisation divides by a range of ~20)."""
import numpy as np
import pytest
import torch

import nn_conformer_for_speech_recognition_b200 as b2c
from nn_conformer_for_speech_recognition_b200.frontend import mel_filterbank

gpu = pytest.mark.gpu


def test_mel_filterbank_matches_the_slaney_bank():
    """CPU: the host-built filterbank == torchaudio's slaney/slaney bank (= librosa.filters.mel defaults)."""
    ta = pytest.importorskip("torchaudio")
    for sr, n_fft, n_mels in [(16000, 2048, 40), (16000, 512, 80), (22050, 2048, 128)]:
        fb = mel_filterbank(sr, n_fft, n_mels)
        ref = ta.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, sr / 2.0, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
        assert fb.shape == ref.shape and np.abs(fb - ref).max() <= 1e-5 * ref.max()      # torchaudio builds its bank in fp32


def _reference_logmel(wave, sr=16000, n_mels=40, n_fft=2048, hop=512):
    import torchaudio
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=hop, n_mels=n_mels, f_min=0.0, f_max=sr / 2.0,
                                              power=2.0, center=True, pad_mode="constant", norm="slaney", mel_scale="slaney").double()
    mel = ms(torch.as_tensor(wave, dtype=torch.float64)).numpy()
    mel = np.where(mel < 1e-10, 0, np.log(np.maximum(mel, 1e-300)))       # speechcommands.py:114
    mel = mel - np.min(mel)                                                # :117
    return mel / np.max(mel)                                               # :118-119


@gpu
def test_logmel_frontend_matches_the_reference_steps():
    g = np.random.default_rng(0)
    sr = 16000
    waves = []
    for n in (16000, 12345, 8000, 16000 + 333, 700):                       # 1-s SpeechCommands clips and ragged ones
        t = np.arange(n) / sr
        w = 0.3 * np.sin(2 * np.pi * 440 * t) + 0.1 * np.sin(2 * np.pi * 3000 * t + 1.0) + 0.05 * g.standard_normal(n)
        waves.append(w.astype(np.float32))
    fe = b2c.LogMelFrontend(sr=sr, n_mels=40, device="cuda")
    mels, frames = fe(waves, batch_size=8, max_len=40)
    assert mels.shape == (8, 40, 40) and frames.tolist() == [n // 512 + 1 if n // 512 + 1 <= 40 else 40 for n in (16000, 12345, 8000, 16333, 700)] + [0, 0, 0]
    out = mels.cpu().numpy()
    for i, w in enumerate(waves):
        ref = _reference_logmel(w)
        T = ref.shape[1]
        assert T == len(w) // 512 + 1
        d = np.abs(out[i, :, :T] - ref).max()
        assert d <= 2e-4, f"utterance {i}: max |d| {d:.3e}"
        assert np.all(out[i, :, T:] == 0)                                   # zero padding to max_len
        assert abs(out[i, :, :T].min()) <= 1e-6 and abs(out[i, :, :T].max() - 1) <= 1e-6
    assert np.all(out[5:] == 0)                                            # batch-padding rows


@gpu
def test_padded_batcher_reproduces_get_batch_padding():
    g = torch.Generator().manual_seed(1)
    feats = [torch.rand(40, n, generator=g) for n in (40, 33, 7, 1)]
    pb = b2c.PaddedBatcher(batch_size=6, n_features=40, max_len=40, device="cuda")
    mels, tau = pb(feats)
    assert mels.shape == (6, 40, 40) and tau.tolist() == [40, 33, 7, 1, 0, 0]
    ref = torch.zeros(6, 40, 40)
    for i, f in enumerate(feats):
        ref[i, :, :f.shape[1]] = f                                         # speechcommands.py:186-188: pad rows, then the batch
    assert torch.equal(mels.cpu(), ref)
    _, tau2 = pb(feats, tau=[40, 40, 40, 40])                              # the reference's tau = len(mel) quirk (speechcommands.py:71)
    assert tau2.tolist() == [40, 40, 40, 40, 0, 0]
    with pytest.raises(ValueError):
        pb([torch.rand(40, 41)])
