"""CPU-only checks of the host side: drop-in interfaces (names, state_dict keys/shapes, initialisation identical to the
reference's modules), SpecAugment draw order against the oracle, no silent fallback, no oracle use in the product."""
import os
import random
import re
import sys

import numpy as np
import pytest
import torch

import nn_conformer_for_speech_recognition_b200 as b2c
from nn_conformer_for_speech_recognition_b200 import specaugment
from oracle import specaugment_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nn_conformer_for_speech_recognition_b200")
REF_STAGED = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "lib"))


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f"{f} imports oracle/"
                assert "oracle/" not in text and "oracle." not in text.replace("the oracle.", ""), f"{f} refers to oracle/"


@pytest.mark.parametrize("args", [(144, 4, 576, 2, 31, 0.1), (256, 4, 1024, 2, 31, 0.5), (512, 8, 512, 1, 33, 0.5)])
def test_conformer_state_dict_and_init_match_torchaudio(args):
    ta = pytest.importorskip("torchaudio")
    torch.manual_seed(3)
    ours = b2c.Conformer(*args)
    torch.manual_seed(3)
    ref = ta.models.Conformer(*args)
    sa, sb = ours.state_dict(), ref.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k
    ref.load_state_dict(sa)                       # checkpoints move in both directions
    ours.load_state_dict(sb)
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]


def test_conformer_ctor_errors_like_torchaudio():
    with pytest.raises(ValueError):
        b2c.Conformer(64, 4, 128, 1, 32)           # even depthwise kernel (torchaudio conformer.py:40-41)
    with pytest.raises(AssertionError):
        b2c.Conformer(66, 4, 128, 1, 31)           # embed_dim not divisible by heads (nn.MultiheadAttention)


def test_no_cpu_fallback():
    m = b2c.Conformer(64, 4, 128, 1, 31)
    with pytest.raises(b2c.B2CError):
        m(torch.zeros(1, 8, 64), torch.tensor([8]))
    if not torch.cuda.is_available():
        from nn_conformer_for_speech_recognition_b200 import ops
        with pytest.raises((b2c.B2CError, ValueError)):
            ops.layernorm_fwd(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))


class _HP:
    conv_sub_1_nodes, conv_sub_1_kernel, conv_sub_1_stride = 512, 7, (2, 2)
    conv_sub_2_kernel, conv_sub_2_stride = 3, (2, 2)
    input_rows, input_cols = 80, 1500


def test_convsub_interface_and_out_size():
    m = b2c.ConvSubSampling(_HP(), 1, 128)
    assert m.out_size == 128 * 18 * 373             # (n - k + s)//s twice: 80->37->18, 1500->747->373
    assert list(m.state_dict().keys()) == ["conv_sub_1.weight", "conv_sub_1.bias", "conv_sub_2.weight", "conv_sub_2.bias"]
    assert m.state_dict()["conv_sub_1.weight"].shape == (512, 1, 7, 7)
    assert m.state_dict()["conv_sub_2.weight"].shape == (128, 512, 3, 3)
    assert (m.kc, m.stride, m.KP) == (11, 4, 128)
    with pytest.raises(NotImplementedError):
        b2c.ConvSubSampling(_HP(), 3, 128)           # in_size != 1 is not on the reference's path


@pytest.mark.skipif(not REF_STAGED, reason="oracle/_ref not staged")
def test_convsub_matches_staged_reference_module():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    from lib.convsubsampling import ConvSubSampling as RefCS
    torch.manual_seed(1)
    r = RefCS(_HP(), 1, 128)
    torch.manual_seed(1)
    m = b2c.ConvSubSampling(_HP(), 1, 128)
    assert m.out_size == r.out_size
    for (ka, va), (kb, vb) in zip(m.state_dict().items(), r.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


class _Model:
    def __init__(self, **kw):
        class HP:
            pass
        self.hp = HP()
        for k, v in {**so.hp_defaults(), **kw}.items():
            setattr(self.hp, k, v)


@pytest.mark.parametrize("kw,tau", [({}, [12, 9, 0]), ({}, [40, 40, 33, 3, 0]),
                                    (dict(adaptive_size=True, warping_param_W=3, warping_ntimes=2, time_multiplicity=3),
                                     [64, 50, 41, 7]),
                                    (dict(adaptive_multiplicity=True), [30, 30])])
def test_specaug_draw_order_matches_oracle(kw, tau):
    """The host half of SpecAugment (which draws, in which order, with which bounds) against the oracle, and the
    global `random` stream ends in the same state."""
    model = _Model(**kw)
    hp = {**so.hp_defaults(), **kw}
    B, F = len(tau), 20
    random.seed(99)
    warps = [specaugment._draw_warp(model.hp, B, tau) for _ in range(hp["warping_ntimes"])]
    fm = [specaugment._draw_fmask(model.hp, F) for _ in range(hp["frequency_mask_ntimes"])]
    Mt = hp["time_multiplicity"] if not hp["adaptive_multiplicity"] else min(hp["time_multiplicity"], int(hp["pm"]))
    tm = [specaugment._draw_tmask(model.hp, B, tau) for _ in range(Mt)]
    state = random.getstate()
    ref = so.draw_params(so.PyRandom(99), hp, B, F, tau)
    assert [(list(w), list(w0)) for w, w0 in warps] == [(list(w), list(w0)) for w, w0 in ref["warps"]]
    assert fm == ref["fmasks"]
    assert [[t0 for t0, _ in m] for m in tm] == [t0s for t0s, _ in ref["tmasks"]]
    assert [[t for _, t in m] for m in tm] == [ts for _, ts in ref["tmasks"]]
    random.seed(99)
    so.draw_params(random, hp, B, F, tau)
    assert random.getstate() == state


def test_specaug_reference_error_on_tau_equal_2w():
    with pytest.raises(ValueError):                  # randint(W, W-1): asrnn.py:107
        specaugment._draw_warp(_Model().hp, 1, [2])


def test_encoder_output_lengths_formula():
    from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder
    enc = ConformerEncoder(_HP(), 64, 4, 128, 1, 31)
    lens = torch.tensor([1500, 1000, 11, 10, 0])
    assert enc.output_lengths(lens).tolist() == [373, 248, 1, 0, 0]
    assert enc.bridge.weight.shape == (64, 128 * 18)


def test_relative_positional_embeddings_table_is_bit_exact_with_the_reference():
    """SURVEY 8 row a7 (the reference's own piece): `RelativePositionalEmbeddings(shape, relative)()` -- same constructor,
    same call, bit-identical table (fixtures generated by running unused_lib/relativepositionalembeddings.py:17-32,
    tests/golden/make_golden_relpos.py), `requires_grad` following `relative`, half-split sin|cos (not interleaved)."""
    import hashlib
    import numpy as np
    from nn_conformer_for_speech_recognition_b200.relpos import RelativePositionalEmbeddings
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "relpos_table.npz"))
    for shape in [(12, 16), (80, 256), (3, 5, 8)]:
        t = RelativePositionalEmbeddings(shape, True)()
        assert t.requires_grad and tuple(t.shape) == shape
        assert np.array_equal(t.detach().numpy(), g["t_" + "x".join(map(str, shape))])
    big = RelativePositionalEmbeddings((1280, 512), False)()
    assert not big.requires_grad
    assert hashlib.sha256(big.contiguous().numpy().tobytes()).digest() == bytes(g["sha_1280x512"])
    t = RelativePositionalEmbeddings((4, 8))().detach()
    assert torch.equal(t[0], torch.tensor([0., 0., 0., 0., 1., 1., 1., 1.]))          # p = 0: sin half zero, cos half one
    assert abs(float(t[1, 0]) - float(torch.sin(torch.tensor(1.0)))) < 1e-7 and abs(float(t[1, 4]) - float(torch.cos(torch.tensor(1.0)))) < 1e-7


def test_lstm_is_an_nn_lstm_with_the_same_parameters_and_no_cpu_path():
    """b2c.LSTM keeps torch.nn.LSTM's constructor, parameter names / shapes and initialisation (asrnn.py:38), and has no CPU path."""
    torch.manual_seed(11)
    ref = torch.nn.LSTM(256, 512, bidirectional=True)
    torch.manual_seed(11)
    ours = b2c.LSTM(256, 512, bidirectional=True)
    assert isinstance(ours, torch.nn.LSTM)
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys())
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_o[k]), k
    conv = b2c.lstm_from_reference(ref)
    assert all(torch.equal(conv.state_dict()[k], sd_r[k]) for k in sd_r)
    if not torch.cuda.is_available():
        with pytest.raises(b2c.B2CError):
            ours(torch.zeros(3, 2, 256))


def test_prefetch_helpers_refuse_cpu_devices():
    """HostPrefetcher / ResultFetcher are CUDA plumbing (copy stream, pinned slots, events): no CPU path, loud errors."""
    with pytest.raises(ValueError):
        b2c.HostPrefetcher((2, 3), "cpu")
    with pytest.raises(ValueError):
        b2c.ResultFetcher("cpu")


def test_traffic_capture_is_keyed_to_the_kernel_sources():
    """bench.lib_hash() identifies the kernel code by its SOURCES (nvcc builds are not bit-reproducible); the committed ncu
    traffic capture carries the hash of the sources it was taken from, and bench.py ignores it when they differ."""
    import importlib
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    h = bench.lib_hash()
    assert isinstance(h, str) and len(h) == 16 and h == bench.lib_hash()
    cap = json.load(open(os.path.join(root, "profiles", "r02_traffic_train.json")))
    assert "src_sha256_16" in cap and "kernels" in cap
    kernels, note = bench.load_traffic("train", "M")
    assert (len(kernels) > 0) == (cap["src_sha256_16"] == h), note
    assert bench.load_traffic("infer", "M")[0] == {}


def test_layernorm_pair_deferral_never_crosses_a_gradient_exchange_boundary():
    """autograd.defer_first_layernorm: the first LayerNorm of layer li may be left to the pair kernel of layer li-1 only if the
    gradient span that contains it has not been handed to the all-reduce yet (groups of dp.GradSync.attach)."""
    from nn_conformer_for_speech_recognition_b200 import autograd as ag
    L, ng = 16, 4
    bounds = [round(i * L / ng) for i in range(ng + 1)]
    group_lo = {bounds[i]: bounds[i + 1] for i in range(ng)}
    if not ag._LN_PAIR:
        pytest.skip("B2C_LN_PAIR=0")
    for li in range(L):
        single = ag.defer_first_layernorm(li, 256, False, {})
        multi = ag.defer_first_layernorm(li, 256, True, group_lo)
        assert single == (li > 0)                       # layer 0 has no layer below
        assert multi == (li > 0 and li not in group_lo)  # first layer of a group: its span leaves at the end of this layer
        assert not ag.defer_first_layernorm(li, 1024, False, {})
