"""Dropout-on training (the benchmark configuration, hp.dropout = 0.5, lib/hparams.py:43) pinned against an fp64 reference.

torch's Philox dropout stream cannot be reproduced by the kernels, so these tests use MASK REPLAY: the
kernels run forward + backward at p > 0, the masks they really drew are exported with the same device functions
(b2c_dropout_export_mask, and the B2C_DROPPED_PREACT markers the FFN epilogue leaves in the saved pre-activation), and an
fp64 restatement of torchaudio's ConformerLayer (tests/_masked_reference.py) re-applies them.  Output AND every gradient must
then agree within the bf16 tolerance -- a wrong seed offset at any of the backward sites (autograd.py) regenerates a
different mask there and fails the gradient check (demonstrated by test_wrong_backward_seed_is_detected).

Tolerance (stated): forward max|d| <= 4e-2, mean|d| <= 6e-3 on the LayerNormed output; gradients cosine >= 0.998 and relative
L2 <= 5e-2 per parameter (dropout at p = 0.5 doubles the kept activations, so bf16 rounding noise is ~1.4x the p = 0 case).
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nn_conformer_for_speech_recognition_b200 as b2c  # noqa: E402
from _masked_reference import layer_forward  # noqa: E402

gpu = pytest.mark.gpu


def _layer_params(sd, li, dtype=torch.float64):
    pre = f"conformer_layers.{li}."
    return {k[len(pre):]: (v.detach().to("cpu", dtype) if v.is_floating_point() else v.detach().cpu()).clone()
            for k, v in sd.items() if k.startswith(pre)}


def test_masked_reference_equals_torchaudio_without_dropout():
    """CPU: the fp64 restatement used by the mask-replay tests IS torchaudio's layer when no mask is given (train-mode BN)."""
    ta = pytest.importorskip("torchaudio")
    torch.manual_seed(3)
    d, h, f, L, k, B, T = 32, 4, 64, 2, 7, 3, 21
    lens = [21, 13, 5]
    ref = ta.models.Conformer(d, h, f, L, k, 0.0).double().train()
    x = torch.randn(B, T, d, dtype=torch.float64)
    sd = {kk: v.clone() for kk, v in ref.state_dict().items()}
    y_ta, _ = ref(x, torch.tensor(lens))
    y = x
    for li in range(L):
        y = layer_forward(y, _layer_params(sd, li), lens, {}, h)
    assert (y - y_ta).abs().max().item() < 1e-10


def _export(kind, p, seed, rows, cols, epoch):
    from nn_conformer_for_speech_recognition_b200 import _lib
    out = torch.empty((rows, cols), device="cuda", dtype=torch.float32)
    with _lib.epoch_scope(epoch):
        _lib.call("b2c_dropout_export_mask", kind, float(p), int(seed), rows, cols, out.data_ptr(), _lib.stream_ptr())
    return out


def _exported_masks(dbg, li, B, T, d, f, H):
    """scale tensors of layer li, by site (seed offsets as documented in conformer.run_layers / autograd._backward)."""
    from nn_conformer_for_speech_recognition_b200 import _lib
    p, ep = dbg["drop_p"], dbg["epoch"]
    sd = dbg["seed"] + li * 16
    st = dbg["state"][li]
    keep = float(_lib.load().b2c_dropout_keep_scale(p))
    M = B * T

    def hidden(pre):
        u = st[pre]["u"].float()
        return torch.where(u <= -9000.0, torch.zeros_like(u), torch.full_like(u, keep)).view(B, T, f)
    masks = {
        "ffn1_h": hidden("ffn1"), "ffn1_o": _export(0, p, sd + 1, M, d, ep).view(B, T, d),
        "attn": _export(1, p, sd + 4, B * H * T, T, ep).view(B, H, T, T),
        "attn_o": _export(0, p, sd + 5, M, d, ep).view(B, T, d),
        "conv_o": _export(0, p, sd + 8, M, d, ep).view(B, T, d),
        "ffn2_h": hidden("ffn2"), "ffn2_o": _export(0, p, sd + 13, M, d, ep).view(B, T, d),
    }
    return {k: v.double().cpu() for k, v in masks.items()}


def _cmp_grad(got, ref, name, cos_tol, rel_tol, scale):
    g, r = got.double().cpu().ravel().numpy(), ref.double().ravel().numpy()
    nr = np.linalg.norm(r)
    if nr < 1e-7 * scale:                 # analytically ~zero (depthwise bias under train-mode BatchNorm)
        assert np.linalg.norm(g) <= rel_tol * scale, f"{name}: reference ~0, got {np.linalg.norm(g):.3e}"
        return 1.0, 0.0
    cos = float(g @ r / (np.linalg.norm(g) * nr + 1e-300))
    rel = float(np.linalg.norm(g - r) / nr)
    assert cos >= cos_tol and rel <= rel_tol, f"{name}: cosine {cos:.5f} rel-L2 {rel:.3e}"
    return cos, rel


def _run_replay(p, d, h, f, L, k, B, T, lens, seed_shift=0):
    torch.manual_seed(11)
    m = b2c.Conformer(d, h, f, L, k, p).cuda().train()
    m._debug_keep_state = True
    x = torch.randn(B, T, d)
    w = torch.randn(B, T, d, dtype=torch.float64) * (torch.arange(T)[None, :, None] < torch.tensor(lens)[:, None, None])
    sd0 = {kk: v.detach().cpu().clone() for kk, v in m.state_dict().items()}
    xg = x.cuda().requires_grad_(True)
    y, _ = m(xg, torch.tensor(lens).cuda())
    (y.double() * w.cuda()).sum().backward()
    dbg = m._debug_state
    assert dbg["drop_p"] == p and dbg["epoch"] is not None
    dbg = dict(dbg, seed=dbg["seed"] + seed_shift)
    # fp64 reference with the exported masks
    xr = x.double().requires_grad_(True)
    params, yr = [], xr
    for li in range(L):
        P = {kk: (v.requires_grad_(True) if v.is_floating_point() and "running" not in kk else v)
             for kk, v in _layer_params(sd0, li).items()}
        params.append(P)
        yr = layer_forward(yr, P, lens, _exported_masks(dbg, li, B, T, d, f, h), h)
    (yr * w).sum().backward()
    return m, y, xg, yr, xr, params


@gpu
@pytest.mark.parametrize("p,d,h,f,L,k,B,T,lens", [
    (0.5, 128, 2, 256, 2, 31, 2, 150, [150, 97]),          # the reference's dropout (asrnn.py:29), one-bit mask paths
    (0.5, 256, 4, 1024, 1, 31, 2, 373, [373, 200]),        # Conformer-M layer geometry, 15 s
    (0.25, 128, 2, 256, 1, 31, 2, 140, [140, 140]),        # general-p paths (8-bit decisions)
])
def test_dropout_masks_replayed_in_fp64_reference(p, d, h, f, L, k, B, T, lens):
    m, y, xg, yr, xr, params = _run_replay(p, d, h, f, L, k, B, T, lens)
    valid = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None])
    diff = (y.detach().cpu().double() - yr.detach())[valid].abs()
    assert diff.max().item() <= 4e-2 and diff.mean().item() <= 6e-3, f"forward: max {diff.max():.3e} mean {diff.mean():.3e}"
    _cmp_grad(xg.grad, xr.grad, "d loss / d x", 0.998, 5e-2, 1.0)
    worst = (1.0, 0.0, "")
    for li in range(L):
        scale = float(params[li]["conv_module.sequential.2.weight"].grad.norm())
        for name, pr in params[li].items():
            if not (pr.is_floating_point() and pr.requires_grad):
                continue
            got = dict(m.named_parameters())[f"conformer_layers.{li}.{name}"].grad
            cos, rel = _cmp_grad(got, pr.grad, f"layer {li} {name}", 0.998, 5e-2, scale)
            if rel > worst[1]:
                worst = (cos, rel, f"layer {li} {name}")
    print(f"dropout p={p}: worst gradient {worst[2]}: cosine {worst[0]:.5f} rel-L2 {worst[1]:.3e}")


@gpu
def test_wrong_backward_seed_is_detected():
    """Sanity of the method: replaying masks drawn with a DIFFERENT seed must break gradient parity (so a wrong site
    offset in the backward pass cannot hide behind the tolerance)."""
    with pytest.raises(AssertionError):
        m, y, xg, yr, xr, params = _run_replay(0.5, 128, 2, 256, 1, 31, 2, 150, [150, 97], seed_shift=1)
        _cmp_grad(xg.grad, xr.grad, "d loss / d x", 0.998, 5e-2, 1.0)


@gpu
def test_two_forwards_then_two_backwards_use_their_own_masks():
    """Advisor finding (round 1): with ONE global epoch word, backward(A) after forward(B) regenerated B's masks.  Each
    training forward now snapshots its own epoch word: (fwd A, fwd B, bwd B, bwd A) gives A the gradients of (fwd A, bwd A)."""
    from nn_conformer_for_speech_recognition_b200 import _lib
    torch.manual_seed(5)
    d, h, f, L, k, B, T = 128, 2, 256, 2, 31, 2, 130
    m = b2c.Conformer(d, h, f, L, k, 0.5).cuda().train()
    lens = torch.tensor([130, 71]).cuda()
    xa, xb = torch.randn(B, T, d).cuda(), torch.randn(B, T, d).cuda()

    def reset():
        _lib.call("b2c_dropout_epoch_set", 100, _lib.stream_ptr())
        m._seed_counter = 0
        m.zero_grad(set_to_none=True)

    reset()
    xa1 = xa.clone().requires_grad_(True)
    ya1, _ = m(xa1, lens)
    ya1.square().mean().backward()
    g_single = {n: p.grad.clone() for n, p in m.named_parameters()}
    gx_single = xa1.grad.clone()

    reset()
    xa2 = xa.clone().requires_grad_(True)
    ya2, _ = m(xa2, lens)                       # epoch 101, seed #1  (same as above)
    yb, _ = m(xb.clone().requires_grad_(True), lens)   # epoch 102, seed #2
    yb.square().mean().backward()
    m.zero_grad(set_to_none=True)               # keep only A's parameter gradients
    ya2.square().mean().backward()
    dy = (ya1 - ya2).abs()
    assert dy.max().item() <= 5e-3 and dy.mean().item() <= 2e-5        # same masks (BatchNorm's fp64 atomics may reorder)
    # identical masks -> identical math up to the order of fp32 atomics in the split-K weight gradients
    assert (xa2.grad - gx_single).abs().max().item() <= 2e-3 * gx_single.abs().max().item() + 1e-9
    for n, p in m.named_parameters():
        ref = g_single[n]
        err = (p.grad - ref).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item() + 1e-7, f"{n}: {err:.3e} vs scale {ref.abs().max().item():.3e}"


@gpu
def test_train_mode_without_autograd_still_applies_dropout():
    """torchaudio applies dropout in model.train() under torch.no_grad() (MC-dropout / noisy-student use); so do we, with
    the same masks the autograd path draws for the same (seed, epoch)."""
    from nn_conformer_for_speech_recognition_b200 import _lib
    torch.manual_seed(6)
    m = b2c.Conformer(128, 2, 256, 1, 31, 0.5).cuda().train()
    lens = torch.tensor([140, 99]).cuda()
    x = torch.randn(2, 140, 128).cuda()

    def reset():
        _lib.call("b2c_dropout_epoch_set", 7, _lib.stream_ptr())
        m._seed_counter = 0
    reset()
    with torch.no_grad():
        y_ng, _ = m(x, lens)
    reset()
    y_g, _ = m(x.clone().requires_grad_(True), lens)
    m.dropout_p = 0.0                                           # same train-mode BatchNorm, dropout off
    with torch.no_grad():
        y_nodrop, _ = m(x, lens)
    assert (y_ng - y_g.detach()).abs().max().item() <= 5e-3
    assert (y_ng - y_nodrop).abs().mean().item() > 0.05         # dropout really was applied


@gpu
def test_reference_positional_table_added_in_front_of_the_attention_layernorm():
    """The reference's own positional variant (unused_lib/conformer.py:101-103): `x + RelativePositionalEmbeddings(...)()` in
    front of the attention LayerNorm, one table row per flattened token.  `Conformer(..., pos_table=E)` vs the fp64 layer
    restatement with the same addition: forward and all gradients (bf16 tolerance of the module tests)."""
    torch.manual_seed(21)
    d, h, f, L, k, B, T = 128, 2, 256, 2, 31, 2, 140
    lens = [140, 101]
    E = b2c.RelativePositionalEmbeddings((B * T, d), True)().detach()
    m = b2c.Conformer(d, h, f, L, k, 0.0, pos_table=E).cuda().train()
    assert "pos_table" not in m.state_dict()
    x = torch.randn(B, T, d)
    w = torch.randn(B, T, d, dtype=torch.float64) * (torch.arange(T)[None, :, None] < torch.tensor(lens)[:, None, None])
    sd0 = {kk: v.detach().cpu().clone() for kk, v in m.state_dict().items()}
    xg = x.cuda().requires_grad_(True)
    y, _ = m(xg, torch.tensor(lens).cuda())
    (y.double() * w.cuda()).sum().backward()
    xr = x.double().requires_grad_(True)
    params, yr = [], xr
    for li in range(L):
        P = {kk: (v.requires_grad_(True) if v.is_floating_point() and "running" not in kk else v)
             for kk, v in _layer_params(sd0, li).items()}
        params.append(P)
        yr = layer_forward(yr, P, lens, {}, h, pos=E.double())
    (yr * w).sum().backward()
    valid = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None])
    diff = (y.detach().cpu().double() - yr.detach())[valid].abs()
    assert diff.max().item() <= 3e-2 and diff.mean().item() <= 5e-3
    # and it is NOT a no-op: without the table the output differs
    m2 = b2c.Conformer(d, h, f, L, k, 0.0).cuda().train()
    m2.load_state_dict(m.state_dict())
    y2, _ = m2(x.cuda(), torch.tensor(lens).cuda())
    assert (y2 - y).abs().mean().item() > 0.05
    _cmp_grad(xg.grad, xr.grad, "d loss / d x", 0.999, 3e-2, 1.0)
    for li in range(L):
        scale = float(params[li]["conv_module.sequential.2.weight"].grad.norm())
        for name, pr in params[li].items():
            if pr.is_floating_point() and pr.requires_grad:
                _cmp_grad(dict(m.named_parameters())[f"conformer_layers.{li}.{name}"].grad, pr.grad, f"layer {li} {name}", 0.999, 3e-2, scale)
    # eval path too (per-launch pipeline; the one-call-per-layer entry point has no table argument)
    m.eval()
    with torch.no_grad():
        ye, _ = m(x.cuda(), torch.tensor(lens).cuda())
    assert torch.isfinite(ye).all()
