"""End-to-end parity of the BENCHMARKED path (bench.py's training step) on a real B200.

    ours:       SpecAugPlan -> ConvSubSampling (composed 11x11/s4 conv) -> bridge GEMM (permuted weight) -> 16-layer
                Conformer stack (one autograd node) -> b2c.Linear head -> log_softmax + CTC
    reference:  oracle/ref_pipeline.RefEncoder = the reference's own ConvSubSampling + SpecAugment (staged copy under
                oracle/_ref), nn.Linear bridge in the reference's NCHW flatten order, torchaudio.models.Conformer, nn.Linear
                head, torch CTCLoss(blank=0, zero_infinity=True)  (asrnn.py:193-221,254-256; runner.py:140-146), fp64 CPU

One state_dict is loaded into both.  Compared: encoder output, CTC loss and EVERY parameter gradient (conv_sub_1/2 at the
real 512 -> 128 channels, bridge, all 16 layers, head), in reference- and paper-mode SpecAugment under the same
`random.seed`; then the same step through torch.cuda.make_graphed_callables replay; then N graphed optimizer steps
against N eager ones (the stale-weight bug the round-1 advisor found: weight preparation must be part of the graph).

Stated tolerance at full depth (L = 16; every activation gradient crosses ~10 bf16 roundings per layer on its way down):
forward max|d| <= 4e-2 / mean|d| <= 6e-3 on the LayerNormed encoder output; CTC loss relative 2e-3; gradients cosine >= 0.995
and relative L2 <= 0.1 per parameter (measured worst case is printed).
"""
import os
import random
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import nn_conformer_for_speech_recognition_b200 as b2c  # noqa: E402
from nn_conformer_for_speech_recognition_b200.encoder import ConformerEncoder  # noqa: E402
from nn_conformer_for_speech_recognition_b200.specaugment import SpecAugPlan  # noqa: E402
from oracle import ref_pipeline as rp  # noqa: E402

CFG_M = dict(d=256, h=4, f=1024, L=16, k=31)


def _pair(cfg, T, dropout=0.0, dtype=torch.float64, ntok=rp.NTOK):
    """(reference pipeline on CPU, ours on the GPU) sharing one state_dict."""
    torch.manual_seed(0)
    ref = rp.RefEncoder(cfg, T, dropout=dropout, ntok=ntok)
    enc = ConformerEncoder(rp.make_hparams(80, T), cfg["d"], cfg["h"], cfg["f"], cfg["L"], cfg["k"], dropout=dropout,
                           specaug_mode="reference")
    head = b2c.Linear(cfg["d"], ntok)
    sd = ref.state_dict()
    enc.load_state_dict({k: v for k, v in sd.items() if not k.startswith("head.")})
    head.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("head.")})
    return ref.to(dtype), enc.cuda(), head.cuda()


def _batch(B, T, lens, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 1, 80, T, generator=g)
    tgt = torch.randint(1, rp.NTOK, (B, rp.TGT_LEN), generator=g)
    return x, torch.tensor(lens), tgt, torch.full((B,), rp.TGT_LEN)


def _grad_stats(ours, ref):
    out = {}
    for name, g_ref in ref.items():
        g = ours[name].double().cpu().ravel().numpy()
        r = g_ref.double().ravel().numpy()
        nr = np.linalg.norm(r)
        if nr < 1e-12:
            continue
        out[name] = (float(g @ r / (np.linalg.norm(g) * nr + 1e-300)), float(np.linalg.norm(g - r) / nr))
    return out


@pytest.mark.parametrize("mode", ["reference", "paper"])
def test_benchmarked_training_step_vs_reference_pipeline(mode):
    """Yardstick: the synthetic benchmark input (U[0,1) log-mel, random-init weights) is badly conditioned for ANY bf16
    evaluation -- the bridge output is a common-mode vector (RMS 0.096) plus token-to-token variation of RMS 0.056, and the
    16 train-mode BatchNorms renormalise that variation.  torch's own CPU bf16 autocast of the reference modules deviates
    from fp64 by max 1.7e-1 / mean 2.7e-2 on this input (fp32: 1.4e-5).  The bar is therefore stated RELATIVE to that:
    our error vs fp64 must not exceed the error of torch's bf16 autocast of the reference pipeline, output and gradients."""
    B, T = 4, 1500
    lens = [1500, 1500, 1337, 1111]
    ref, enc, head = _pair(CFG_M, T)
    x, lens_t, tgt, tgt_lens = _batch(B, T, lens)
    # ---- reference: fp64 CPU, the reference's own modules ----
    ref.train()
    random.seed(77)
    y_r, l2 = ref.encode(x.double(), lens_t, specaug=mode)
    logits_r = ref.head(y_r)
    state_after = random.getstate()
    loss_r = rp.ctc_loss(logits_r, tgt, l2, tgt_lens)
    loss_r.backward()
    g_ref = {n: p.grad.clone() for n, p in ref.named_parameters()}
    # ---- yardstick: the same modules under torch's bf16 autocast (fp32 parameters) ----
    ref.float().zero_grad(set_to_none=True)
    random.seed(77)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        y_a, _ = ref.encode(x, lens_t, specaug=mode)
        logits_a = ref.head(y_a)
    loss_a = rp.ctc_loss(logits_a.float(), tgt, l2, tgt_lens)
    loss_a.backward()
    g_ac = {n: p.grad.clone() for n, p in ref.named_parameters()}
    # ---- ours ----
    enc.train(); head.train()
    enc._b2c_specaug_mode = mode
    random.seed(77)
    y, out_lens = enc(x.cuda(), lens_t.cuda(), SpecAugment=True, lens_host=lens)
    assert random.getstate() == state_after, "SpecAugment consumed a different number of draws than the reference"
    assert out_lens.tolist() == l2.tolist()
    logits = head(y)
    loss = rp.ctc_loss(logits, tgt.cuda(), out_lens, tgt_lens.cuda())
    loss.backward()
    valid = (torch.arange(y.shape[1])[None, :] < l2[:, None])
    d = (y.detach().cpu().double() - y_r.detach())[valid].abs()
    da = (y_a.detach().double() - y_r.detach())[valid].abs()
    rel_loss = abs(loss.item() - loss_r.item()) / abs(loss_r.item())
    rel_loss_a = abs(loss_a.item() - loss_r.item()) / abs(loss_r.item())
    print(f"[{mode}] encoder output vs fp64: ours max {d.max():.3e} mean {d.mean():.3e} | torch bf16 autocast max {da.max():.3e} "
          f"mean {da.mean():.3e}; CTC loss {loss.item():.5f} (fp64 {loss_r.item():.5f}: rel {rel_loss:.2e}, autocast rel {rel_loss_a:.2e})")
    assert torch.isfinite(y).all()
    assert d.mean().item() <= da.mean().item() and d.max().item() <= 1.25 * da.max().item()
    assert rel_loss <= max(2e-3, 1.5 * rel_loss_a), f"CTC loss {loss.item():.6f} vs {loss_r.item():.6f}"
    ours = {n: p.grad for n, p in enc.named_parameters()}
    ours.update({"head." + n: p.grad for n, p in head.named_parameters()})
    assert set(ours) == set(g_ref)
    so, sa = _grad_stats(ours, g_ref), _grad_stats(g_ac, g_ref)
    rows = sorted(((so[n][1], so[n][0], sa[n][1], sa[n][0], n) for n in so), reverse=True)
    print(f"[{mode}] worst gradients (rel-L2 ours, cos ours | rel-L2 autocast, cos autocast):")
    for r in rows[:6]:
        print(f"    {r[0]:.3e} {r[1]:.5f} | {r[2]:.3e} {r[3]:.5f}  {r[4]}")
    med_o, med_a = float(np.median([v[1] for v in so.values()])), float(np.median([v[1] for v in sa.values()]))
    print(f"[{mode}] median rel-L2 over {len(so)} parameters: ours {med_o:.3e}, autocast {med_a:.3e}")
    bad = []
    for n in so:
        if "conv_module.sequential.2.bias" in n:      # feeds a train-mode BatchNorm: true gradient ~0, rounding noise only
            continue
        if not (so[n][1] <= max(5e-2, 1.25 * sa[n][1]) and so[n][0] >= min(0.998, sa[n][0] - 2e-3)):
            bad.append((n, so[n], sa[n]))
    assert not bad, f"{len(bad)} parameter gradients are worse than torch's bf16 autocast: {bad[:5]}"
    assert med_o <= med_a


def test_graph_replay_equals_eager_step():
    """The path bench.py times: encoder forward + backward captured by make_graphed_callables, SpecAugPlan parameters
    drawn on the host before each replay.  Two replays with fresh draws must reproduce the eager results."""
    B, T = 4, 1500
    lens = [1500, 1400, 1300, 900]
    _, enc, head = _pair(CFG_M, T)
    enc.train(); head.train()
    enc._b2c_specaug_mode = "paper"
    x, lens_t, tgt, tgt_lens = _batch(B, T, lens)
    xd, lens_d = x.cuda(), lens_t.cuda()
    plan = SpecAugPlan(enc.hp, B, "cuda", mode="paper")

    class Step(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = enc

        def forward(self, xin):
            return self.enc(xin, lens_d, SpecAugment=True, specaug_plan=plan)[0]

    def run(fn, seed):
        random.seed(seed)
        plan.draw(80, lens)
        enc.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
        y = fn(xd)
        loss = rp.ctc_loss(head(y), tgt.cuda(), enc.output_lengths(lens_d), tgt_lens.cuda())
        loss.backward()
        return loss.item(), y.detach().clone(), {n: p.grad.clone() for n, p in enc.named_parameters()}

    eager = Step()
    e1, e1b, e2 = run(eager, 1), run(eager, 1), run(eager, 2)
    random.seed(0)
    plan.draw(80, lens)
    graphed = torch.cuda.make_graphed_callables(Step(), (xd,), num_warmup_iters=3)
    g1, g2 = run(graphed, 1), run(graphed, 2)
    assert abs(e1[0] - e2[0]) > 0, "the two SpecAugment draws should differ"
    # Forward: the same kernels on the same operands -> bit-identical.  Backward: parameter gradients are accumulated with
    # fp32 atomics (split-K weight gradients, column sums), whose order varies from run to run; on this badly conditioned
    # synthetic problem (see the test above) those last-bit differences are amplified on the way down 16 layers, so two
    # EAGER runs already differ by up to ~1 % of a gradient's scale.  Graph-vs-eager must stay within that run-to-run
    # noise (10x the eager-vs-eager difference, floor 2e-3, cap 5e-2 of the scale).
    for (le, ye, ge), (lg, yg, gg) in ((e1, g1), (e2, g2)):
        assert torch.equal(ye, yg), f"encoder output differs: max {(ye - yg).abs().max().item():.3e}"
        assert abs(le - lg) <= 1e-6 * abs(le)
    worst = 0.0
    for n in e1[2]:
        if n.endswith("conv_module.sequential.2.bias"):    # true gradient ~0 (feeds a train-mode BatchNorm): pure rounding noise
            continue
        sc = e1[2][n].abs().max().item()
        noise = (e1[2][n] - e1b[2][n]).abs().max().item()
        err = (e1[2][n] - g1[2][n]).abs().max().item()
        worst = max(worst, err / (sc + 1e-30))
        assert err <= min(max(10 * noise, 2e-3 * sc), 5e-2 * sc) + 1e-8, f"{n}: graph vs eager {err:.3e}, eager vs eager {noise:.3e} (scale {sc:.3e})"
    print(f"graph vs eager: worst parameter-gradient difference {worst:.2e} of its scale")


def test_graphed_optimizer_steps_match_eager_steps():
    """N graph-replayed training steps with the fused Adafactor == N eager steps: the bf16 operand copies, the composed
    conv-subsampling weights and the permuted bridge weight must be RE-DERIVED from the fp32 masters inside the graph
    (round-1 advisor: they were prepared on the host only, so the graph trained against the initial weights forever)."""
    from nn_conformer_for_speech_recognition_b200.optim import Adafactor
    cfg = dict(d=256, h=4, f=1024, L=2, k=31)
    B, T, N_STEPS, LR = 4, 600, 4, 2e-3
    lens = [600, 600, 555, 480]
    x, lens_t, tgt, tgt_lens = _batch(B, T, lens)
    xd, lens_d = x.cuda(), lens_t.cuda()

    def make():
        _, enc, head = _pair(cfg, T)
        enc.train(); head.train()
        enc._b2c_specaug_mode = "paper"
        return enc, head

    def train(graph):
        enc, head = make()
        plan = SpecAugPlan(enc.hp, B, "cuda", mode="paper")
        opt = Adafactor(list(enc.parameters()) + list(head.parameters()), lr=LR, beta1=0.9)

        class Step(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.enc = enc

            def forward(self, xin):
                return self.enc(xin, lens_d, SpecAugment=True, specaug_plan=plan)[0]
        fn = Step()
        if graph:
            random.seed(0)
            plan.draw(80, lens)
            fn = torch.cuda.make_graphed_callables(fn, (xd,), num_warmup_iters=3)
            for m in enc.modules():                    # warm-up ran train-mode BatchNorm: restore the buffers
                if isinstance(m, torch.nn.BatchNorm1d):
                    m.reset_running_stats()
        losses = []
        for s in range(N_STEPS):
            random.seed(100 + s)
            plan.draw(80, lens)
            opt.zero_grad(set_to_none=True)
            y = fn(xd)
            loss = rp.ctc_loss(head(y), tgt.cuda(), enc.output_lengths(lens_d), tgt_lens.cuda())
            loss.backward()
            opt.step()
            losses.append(loss.item())
        if graph:
            fn(xd)                                     # one more REPLAY (no Python runs inside): must recast the weights
        return enc, head, losses

    enc_e, head_e, loss_e = train(False)
    enc_g, head_g, loss_g = train(True)
    # the operands the graph reads ARE the current parameters (cast / composed / permuted inside the graph)
    w = enc_g.conformers._cache.bf16[0]["ffn1.sequential.1.weight"]
    p = enc_g.conformers.conformer_layers[0].ffn1.sequential[1].weight
    assert torch.equal(w, p.detach().to(torch.bfloat16)), "graph replays left the bf16 weight copy stale"
    print("losses eager", loss_e, "graph", loss_g)
    assert loss_e[0] != loss_e[-1]
    for a, b_ in zip(loss_e, loss_g):
        assert abs(a - b_) <= 2e-3 * abs(a), f"loss trajectories diverge: {loss_e} vs {loss_g}"
    pe, pg = dict(enc_e.named_parameters()), dict(enc_g.named_parameters())
    moved = 0.0
    for n in pe:
        diff = (pe[n] - pg[n]).abs().mean().item()
        assert diff <= 0.15 * LR * N_STEPS + 1e-7, f"{n}: parameters differ after {N_STEPS} steps (mean |d| {diff:.3e})"
        moved = max(moved, diff)
    print(f"max mean-|d| between eager and graphed parameters: {moved:.3e} (lr*steps = {LR * N_STEPS:.1e})")


def test_specaug_plan_matches_specaugment_and_survives_run_ahead():
    """SpecAugPlan.draw/apply == specaugment.SpecAugment under the same seed, and a host that runs ahead of the device
    (no synchronisation between steps) never tears the parameter block (double-buffered pinned slots)."""
    from nn_conformer_for_speech_recognition_b200 import specaugment as sa
    B, F, T = 8, 80, 700
    hp = rp.make_hparams(F, T)
    lens = [700, 650, 600, 500, 400, 300, 10, 0]
    x = torch.rand(B, 1, F, T, generator=torch.Generator().manual_seed(5)).cuda()

    class Holder:
        pass
    holder = Holder()
    holder.hp, holder._b2c_specaug_mode = hp, "paper"
    plan = SpecAugPlan(hp, B, "cuda", mode="paper")
    expected, got = [], []
    for s in range(6):
        random.seed(900 + s)
        expected.append(sa.SpecAugment(holder, x, lens))
    torch.cuda.synchronize()
    big = torch.empty(64 << 20, device="cuda")
    for s in range(6):                                   # no sync inside: the host runs ahead of the queued work
        random.seed(900 + s)
        plan.draw(F, lens)
        for _ in range(20):
            big.normal_()                                # keep the stream busy so copies queue up behind kernels
        got.append(plan.apply(x))
    torch.cuda.synchronize()
    for s in range(6):
        assert torch.equal(expected[s], got[s]), f"step {s}: plan output differs from SpecAugment"
