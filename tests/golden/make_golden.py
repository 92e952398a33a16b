"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in the build container.

Run once here (needs /root/reference staged by oracle/stage_ref.py, torchaudio, transformers):
    python tests/golden/make_golden.py
The fixtures are committed; the GPU box never needs /root/reference.

Sources exercised (all UNMODIFIED reference / third-party code):
  * lib.convsubsampling.ConvSubSampling                  (/root/reference/lib/convsubsampling.py)
  * torchaudio.models.Conformer                          (site-packages/torchaudio/models/conformer.py)
  * lib.standard.asrnn.ASRNN.SpecAugment & friends       (/root/reference/lib/standard/asrnn.py:91-192)
  * HF Wav2Vec2ConformerSelfAttention (relative)         (pinning the rel_shift variant)
"""
import hashlib
import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from oracle import stage_ref  # noqa: E402

assert stage_ref.stage(), "reference not available"
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from lib.convsubsampling import ConvSubSampling  # noqa: E402
from lib.hparams import HParams  # noqa: E402
from lib.standard.asrnn import ASRNN  # noqa: E402
from torchaudio.models import Conformer  # noqa: E402


def fill(module):
    """FILL(module) of SURVEY.md Appendix B."""
    sd = module.state_dict()
    for i, (key, v) in enumerate(sd.items()):
        if not v.is_floating_point():
            continue
        idx = torch.arange(v.numel(), dtype=torch.float64).reshape(v.shape)
        s = torch.sin(0.37 * i + 0.11 * idx)
        if key.endswith("running_var"):
            val = 1 + 0.1 * s
        elif key.endswith("running_mean"):
            val = 0.05 * s
        elif v.dim() == 1 and key.endswith("weight") and ("layer_norm.weight" in key or (
                "ffn" in key and ".sequential.0.weight" in key) or "conv_module.sequential.3.weight" in key):
            val = 1 + 0.1 * s
        else:
            val = 0.1 * s
        v.copy_(val.to(v.dtype))
    module.load_state_dict(sd)


def kat_input(shape):
    n = int(np.prod(shape))
    return (0.5 * torch.cos(0.01 * torch.arange(n, dtype=torch.float64))).reshape(shape)


def make_hp(rows, cols, **kw):
    hp = HParams(tempfile.mkdtemp())
    hp.device = torch.device("cpu")
    hp.set_input_dim(rows, cols) if hasattr(hp, "set_input_dim") else None
    hp.input_rows, hp.input_cols = rows, cols
    for k, v in kw.items():
        setattr(hp, k, v)
    return hp


def conformer_golden():
    """KAT-2 (d=16,h=2,f=32,L=2,k=7) eval + train, plus grads of every parameter."""
    m = Conformer(16, 2, 32, 2, 7, 0.0).double()
    fill(m)
    x = kat_input((2, 12, 16))
    lengths = torch.tensor([12, 9])
    m.eval()
    with torch.no_grad():
        y_eval, l_out = m(x, lengths)
    m.train()
    xg = x.clone().requires_grad_(True)
    y_train, _ = m(xg, lengths)
    wsum = torch.sin(0.03 * torch.arange(y_train.numel(), dtype=torch.float64)).reshape(y_train.shape)
    loss = (y_train * wsum).sum()
    loss.backward()
    out = dict(x=x.numpy(), lengths=lengths.numpy(), y_eval=y_eval.numpy(), y_train=y_train.detach().numpy(),
               loss=np.float64(loss.item()), dx=xg.grad.numpy(), lengths_out=l_out.numpy())
    for k, p in m.named_parameters():
        out["grad." + k] = p.grad.numpy()
    for k, v in m.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["after." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "conformer_kat2.npz"), **out)
    print("KAT-2 eval sum", y_eval.sum().item(), "abs", y_eval.abs().sum().item())
    print("KAT-2 train sum", y_train.sum().item(), "abs", y_train.abs().sum().item(), "loss", loss.item())


def conformer_random_golden():
    """Seeded-random small Conformer (d=32,h=4,f=64,L=2,k=5), ragged lengths incl. a short one; fp64."""
    torch.manual_seed(7)
    m = Conformer(32, 4, 64, 2, 5, 0.0).double()
    # make BN running stats non-trivial
    for k, v in m.state_dict().items():
        if k.endswith("running_mean"):
            v.copy_(0.1 * torch.randn_like(v))
        if k.endswith("running_var"):
            v.copy_(0.5 + torch.rand_like(v))
    x = torch.randn(3, 21, 32, dtype=torch.float64)
    lengths = torch.tensor([21, 13, 1])
    m.eval()
    with torch.no_grad():
        y, _ = m(x, lengths)
    out = dict(x=x.numpy(), lengths=lengths.numpy(), y_eval=y.numpy())
    for k, v in m.state_dict().items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "conformer_rand.npz"), **out)


def convsub_golden():
    """KAT-3 + a seeded-random case at the real kernel sizes (512ch k7 s2 -> 128ch k3 s2 is too big for a
    fixture; channel counts are reduced, kernel/stride geometry is the reference's)."""
    hp = make_hp(20, 30, conv_sub_1_nodes=8)
    m = ConvSubSampling(hp, 1, 4).double()
    fill(m)
    x = kat_input((2, 1, 20, 30))
    with torch.no_grad():
        z = m(x)
    print("KAT-3", tuple(z.shape), m.out_size, z.sum().item(), z.abs().sum().item(), z[0, 0, 0, :3].tolist())
    out = dict(x=x.numpy(), z=z.numpy(), out_size=np.int64(m.out_size))
    for k, v in m.state_dict().items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "convsub_kat3.npz"), **out)

    torch.manual_seed(11)
    hp = make_hp(80, 61, conv_sub_1_nodes=24)
    m = ConvSubSampling(hp, 1, 16).double()
    x = torch.rand(2, 1, 80, 61, dtype=torch.float64)
    xg = x.clone()
    z = m(xg)
    wsum = torch.cos(0.05 * torch.arange(z.numel(), dtype=torch.float64)).reshape(z.shape)
    (z * wsum).sum().backward()
    out = dict(x=x.numpy(), z=z.detach().numpy(), out_size=np.int64(m.out_size))
    for k, v in m.state_dict().items():
        out["sd." + k] = v.numpy()
    for k, p in m.named_parameters():
        out["grad." + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "convsub_rand.npz"), **out)


class _Trace:
    """Wraps random.randint to record (a, b, r)."""
    def __init__(self):
        self.calls = []
        self._orig = random.randint

    def __enter__(self):
        def rec(a, b):
            r = self._orig(a, b)
            self.calls.append((int(a), int(b), int(r)))
            return r
        random.randint = rec
        return self

    def __exit__(self, *a):
        random.randint = self._orig


def specaug_golden():
    """KAT-1 and two further traces of ASRNN.SpecAugment under random.seed(s)."""
    cases = {
        "kat1": dict(shape=(3, 1, 8, 12), tau=[12, 9, 0], seed=1234, hp={}),
        "b5": dict(shape=(5, 1, 40, 40), tau=[40, 40, 33, 3, 0], seed=42, hp={}),
        "adaptive": dict(shape=(4, 1, 20, 64), tau=[64, 50, 41, 7], seed=7,
                         hp=dict(adaptive_size=True, warping_param_W=3, warping_ntimes=2, time_multiplicity=3)),
    }
    out = {}
    for name, c in cases.items():
        B, _, F, T = c["shape"]
        hp = make_hp(F, T, **c["hp"])
        hp.batch_size, hp.max_len = B, T
        hp.ntokens, hp.blank_index = 16, 0
        model = ASRNN.__new__(ASRNN)          # SpecAugment only reads self.hp (asrnn.py:91-192)
        model.hp = hp
        n = int(np.prod(c["shape"]))
        x = ((1 + torch.arange(n, dtype=torch.float32)) / 1000).reshape(c["shape"])
        tau = torch.tensor(c["tau"])
        random.seed(c["seed"])
        with _Trace() as tr:
            y = ASRNN.SpecAugment(model, x, tau)
        out[name + ".x"] = x.numpy()
        out[name + ".tau"] = tau.numpy()
        out[name + ".seed"] = np.int64(c["seed"])
        out[name + ".y"] = y.numpy()
        out[name + ".draws"] = np.array(tr.calls, dtype=np.int64)
        out[name + ".hp"] = np.array([repr(sorted(c["hp"].items()))])
        print(name, "sha", hashlib.sha256(y.numpy().tobytes()).hexdigest()[:16], "draws", len(tr.calls))
    np.savez_compressed(os.path.join(HERE, "specaug.npz"), **out)


def relpos_golden():
    """rel-pos + rel_shift attention pinned to HF (the reference has no such code)."""
    from transformers import Wav2Vec2ConformerConfig
    from transformers.models.wav2vec2_conformer.modeling_wav2vec2_conformer import (
        Wav2Vec2ConformerRelPositionalEmbedding, Wav2Vec2ConformerSelfAttention)
    torch.manual_seed(3)
    d, h, T, B = 32, 4, 10, 2
    cfg = Wav2Vec2ConformerConfig(hidden_size=d, num_attention_heads=h, position_embeddings_type="relative",
                                  attention_dropout=0.0)
    att = Wav2Vec2ConformerSelfAttention(cfg).double().eval()
    with torch.no_grad():
        att.pos_bias_u.copy_(0.2 * torch.randn_like(att.pos_bias_u))
        att.pos_bias_v.copy_(0.2 * torch.randn_like(att.pos_bias_v))
    emb = Wav2Vec2ConformerRelPositionalEmbedding(cfg).double()
    x = torch.randn(B, T, d, dtype=torch.float64)
    lengths = [10, 6]
    mask = torch.zeros(B, 1, 1, T, dtype=torch.float64)
    for b, l in enumerate(lengths):
        mask[b, :, :, l:] = float("-inf")
    with torch.no_grad():
        pe = emb(x)
        y, _ = att(x, attention_mask=mask, relative_position_embeddings=pe)
    sd = att.state_dict()
    out = dict(x=x.numpy(), lengths=np.array(lengths), y=y.numpy(), pe=pe.numpy(),
               in_w=torch.cat([sd["linear_q.weight"], sd["linear_k.weight"], sd["linear_v.weight"]]).numpy(),
               in_b=torch.cat([sd["linear_q.bias"], sd["linear_k.bias"], sd["linear_v.bias"]]).numpy(),
               out_w=sd["linear_out.weight"].numpy(), out_b=sd["linear_out.bias"].numpy(),
               pos_w=sd["linear_pos.weight"].numpy(), pos_u=sd["pos_bias_u"].numpy(), pos_v=sd["pos_bias_v"].numpy())
    np.savez_compressed(os.path.join(HERE, "relpos_hf.npz"), **out)


if __name__ == "__main__":
    conformer_golden()
    conformer_random_golden()
    convsub_golden()
    specaug_golden()
    relpos_golden()
    print("golden fixtures written to", HERE)
