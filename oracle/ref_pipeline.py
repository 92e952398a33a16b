"""The reference's training / inference pipeline at the benchmark shapes, built from the reference's OWN modules.

TEST / BENCH INFRASTRUCTURE ONLY -- imported by tests/ and by bench.py's reference arm / cpu_baseline leg, never by the
product package.

What it assembles (CPU, fp32 or fp64), following `ASRNN.encoder` (/root/reference/lib/standard/asrnn.py:193-221) and the
training step of `Runner.train` (/root/reference/lib/standard/runner.py:140-146):

    SpecAugment (ASRNN.SpecAugment, asrnn.py:170-192)                      staged copy oracle/_ref/lib/standard/asrnn.py
 -> ConvSubSampling(hp, 1, 128)           (lib/convsubsampling.py:16-45)    staged copy oracle/_ref/lib/convsubsampling.py
 -> bridge: per-frame Linear(128*F', d)   (asrnn.py:207-209 at benchmark shapes: the reference's flatten bridge has
                                           O(T*T') parameters and cannot be instantiated at 15 s)
 -> Dropout(hp.dropout)                   (asrnn.py:210)
 -> torchaudio.models.Conformer           (asrnn.py:29,214; site-packages)
 -> head Linear(d, ntokens) -> log_softmax -> CTCLoss(blank=0, zero_infinity=True)   (asrnn.py:254-256, runner.py:35,143)

When oracle/_ref is absent (never on the build container or the GPU box: build() stages it) the two reference classes are
replaced by their restatements (the same two nn.Conv2d; oracle/specaugment_oracle.py) and `kind` says "port".
"""
from __future__ import annotations

import os
import random
import sys
import tempfile

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
NTOK, TGT_LEN = 1024, 20


def staged() -> bool:
    """True when the staged reference copy is importable (puts it on sys.path)."""
    try:
        from oracle import stage_ref
        ok = stage_ref.stage()
    except Exception:
        ok = os.path.isdir(os.path.join(REF_DIR, "lib"))
    if ok and REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return ok


class _HPFallback:
    """ConvSubSampling / SpecAugment fields of lib/hparams.py:43-95 at their defaults (used only when oracle/_ref is absent)."""
    conv_sub_1_nodes, conv_sub_1_kernel, conv_sub_1_stride = 512, 7, (2, 2)
    conv_sub_2_nodes, conv_sub_2_kernel, conv_sub_2_stride = 128, 3, (2, 2)
    warping_param_W, warping_ntimes = 1, 1
    frequency_mask_param_F, frequency_mask_ntimes = 5, 2
    time_multiplicity, pm, ps, time_mask_param_T = 2, 0.05, 0.05, 5
    adaptive_multiplicity = adaptive_size = False
    mask_value = 0
    dropout = 0.5
    rel_att = False


def make_hparams(rows: int, cols: int, device="cpu"):
    """The reference's own `HParams` (lib/hparams.py, unchanged) with the dataset constructor's setters applied
    (speechcommands.py:39-46); falls back to the restated defaults when the staged copy is missing."""
    if staged():
        from lib.hparams import HParams
        hp = HParams(tempfile.mkdtemp(prefix="b2c_hp_"))
        hp.set_input_dim(rows, cols)
        hp.set_blank_index(0)
        hp.set_ntokens(NTOK)
    else:
        hp = _HPFallback()
        hp.input_rows, hp.input_cols = rows, cols
    hp.device = torch.device(device)
    return hp


class RefEncoder(nn.Module):
    """state_dict keys: conv_sub_sampling.*, bridge.*, conformers.*, head.* (the product's ConformerEncoder + Linear head
    use the same names, so one state_dict loads into both)."""

    def __init__(self, cfg: dict, T: int, dropout: float = 0.5, ntok: int = NTOK, rows: int = 80):
        super().__init__()
        import torchaudio
        self.kind = "reference" if staged() else "port"
        self.hp = make_hparams(rows, T)
        if self.kind == "reference":
            from lib.convsubsampling import ConvSubSampling as RefCS
            self.conv_sub_sampling = RefCS(self.hp, 1, 128)
        else:
            self.conv_sub_sampling = nn.Sequential()
            self.conv_sub_sampling.conv_sub_1 = nn.Conv2d(1, 512, 7, stride=2)
            self.conv_sub_sampling.conv_sub_2 = nn.Conv2d(512, 128, 3, stride=2)
        fo = ((rows - 7) // 2 + 1 - 3) // 2 + 1
        self.bridge = nn.Linear(128 * fo, cfg["d"])
        self.dropout = nn.Dropout(dropout)
        self.conformers = torchaudio.models.Conformer(cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], dropout)
        self.head = nn.Linear(cfg["d"], ntok)

    def spec_augment(self, x, lens, mode="reference"):
        """mode "reference": ASRNN.SpecAugment as shipped (masks drawn, never applied); "paper": same draws,
        masks applied (restated in oracle/specaugment_oracle.py -- the reference has no working implementation of it)."""
        if mode == "reference" and self.kind == "reference":
            from lib.standard.asrnn import ASRNN
            for n in ("time_warping", "frequency_masking", "time_masking"):
                if not hasattr(type(self), n):
                    setattr(type(self), n, getattr(ASRNN, n))
            return ASRNN.SpecAugment(self, x, lens).to(x.dtype)
        from oracle import specaugment_oracle as so
        hpd = so.hp_defaults()
        out = so.spec_augment(x.float().numpy(), [int(v) for v in lens], random, hpd, mode=mode)[0]
        return torch.from_numpy(out).to(x.dtype)

    def encode(self, x, lens, specaug=None):
        if specaug:
            x = self.spec_augment(x, lens, specaug)
        if self.kind == "reference":
            z = self.conv_sub_sampling(x)                                      # (B,128,F',T')
        else:
            z = self.conv_sub_sampling.conv_sub_2(self.conv_sub_sampling.conv_sub_1(x))
        z = z.permute(0, 3, 1, 2).flatten(2)                                   # per-frame features, NCHW flatten order c*F'+f'
        h = self.dropout(self.bridge(z))
        l2 = ((lens - 7) // 2 + 1 - 3) // 2 + 1
        y, _ = self.conformers(h, l2)
        return y, l2

    def forward(self, x, lens, specaug=None):
        y, l2 = self.encode(x, lens, specaug)
        return self.head(y), l2


def ctc_loss(logits, tgt, out_lens, tgt_lens):
    """runner.py:35,143: CTCLoss(blank=0, zero_infinity=True) on log_softmax(logits) (asrnn.py:254-256), time-major."""
    return torch.nn.functional.ctc_loss(logits.log_softmax(-1).transpose(0, 1), tgt, out_lens, tgt_lens, blank=0,
                                        zero_infinity=True)
