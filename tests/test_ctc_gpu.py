"""Fused log-softmax + CTC (csrc/ctc.cu) against torch's own `CTCLoss(blank, zero_infinity=True)` on log_softmax(logits) --
the pair the reference runs at asrnn.py:254-256 / runner.py:35,143 -- in fp64 on the CPU.  Index work is exact; the
arithmetic is fp32 log-space: loss relative error <= 1e-5, gradient max |d| <= max(2e-6 + 1e-4 * max|grad|, 2 x the error of
torch's own fp32 CUDA CTC on the same input) (stated)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import nn_conformer_for_speech_recognition_b200 as b2c  # noqa: E402
from nn_conformer_for_speech_recognition_b200.ctc import CTCLoss, ctc_loss_from_logits  # noqa: E402


def _reference(logits, targets, il, tl, blank, zero_infinity=True):
    x = logits.detach().double().cpu().requires_grad_(True)
    lp = x.log_softmax(-1).transpose(0, 1)
    loss = torch.nn.functional.ctc_loss(lp, targets.cpu().long(), torch.as_tensor(il), torch.as_tensor(tl), blank=blank,
                                        reduction="mean", zero_infinity=zero_infinity)
    loss.backward()
    return loss.item(), x.grad


def _check(B, T, V, S, il, tl, blank=0, seed=0, repeat=False, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    logits = (torch.randn(B, T, V, generator=g) * scale).cuda().requires_grad_(True)
    lo = 1 if blank == 0 else 0
    targets = torch.randint(lo, V - (0 if blank == 0 else 1), (B, S), generator=g)
    if repeat:
        targets[:, 1::2] = targets[:, 0::2][:, :targets[:, 1::2].shape[1]]      # l, l, m, m, ...: repeated labels need the blank
    loss = ctc_loss_from_logits(logits, targets.cuda(), il, tl, blank=blank, zero_infinity=True)
    (loss * 1.7).backward()
    ref_loss, ref_grad = _reference(logits, targets, il, tl, blank)
    assert abs(loss.item() - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
    d = (logits.grad.double().cpu() - 1.7 * ref_grad).abs().max().item()
    # yardstick for long lattices: torch's own fp32 CUDA CTC on the same input.  At T = 373, V = 1024 the log-space
    # quantities reach |log alpha| ~ 2600 where one fp32 ulp is 2.4e-4, so ANY fp32 implementation is ~1e-3 relative here.
    xt = logits.detach().clone().requires_grad_(True)
    lt = torch.nn.functional.ctc_loss(xt.log_softmax(-1).transpose(0, 1), targets.cuda().long(), torch.as_tensor(il).cuda(),
                                      torch.as_tensor(tl).cuda(), blank=blank, reduction="mean", zero_infinity=True)
    (lt * 1.7).backward()
    d_torch = (xt.grad.double().cpu() - 1.7 * ref_grad).abs().max().item()
    gmax = (1.7 * ref_grad).abs().max().item()
    assert d <= max(2e-6 + 1e-4 * gmax, 2.0 * d_torch), f"grad max diff {d:.3e} (torch fp32 CUDA: {d_torch:.3e}, max|grad| {gmax:.3e})"
    return loss.item()


def test_ctc_small_ragged_and_repeated_labels():
    _check(4, 50, 32, 7, [50, 41, 33, 20], [7, 5, 3, 1])
    _check(4, 50, 32, 8, [50, 41, 33, 20], [8, 6, 4, 2], repeat=True, seed=1)
    _check(3, 23, 12, 5, [23, 23, 9], [5, 4, 2], blank=11, seed=2)


def test_ctc_infeasible_and_empty_targets_are_zeroed_like_torch():
    # utterance 1: 4 frames for 6 labels -> infinite loss -> zero_infinity zeroes loss and gradient; utterance 2: empty target;
    # utterance 3: zero input frames
    _check(4, 40, 16, 6, [40, 4, 17, 0], [6, 6, 0, 3], seed=3)


def test_ctc_benchmark_shape():
    B, T, V, S = 64, 373, 1024, 20
    _check(B, T, V, S, [T] * B, [S] * B, seed=4, scale=1.0)
    _check(B, T, V, S, [T - (7 * i) % 200 for i in range(B)], [S - i % 5 for i in range(B)], seed=5, scale=1.0)


def test_ctc_module_takes_time_major_log_probs_like_the_reference():
    g = torch.Generator().manual_seed(9)
    B, T, V, S = 5, 60, 40, 9
    logits = torch.randn(B, T, V, generator=g).cuda()
    lp = logits.log_softmax(-1).requires_grad_(True)
    targets = torch.randint(1, V, (B, S), generator=g)
    il, tl = torch.tensor([60, 55, 50, 45, 40]), torch.tensor([9, 8, 7, 6, 5])
    ours = CTCLoss(blank=0, zero_infinity=True)(lp.transpose(0, 1), targets.cuda(), il, tl)      # runner.py:143 call shape
    ours.backward()
    lp_ref = lp.detach().double().cpu().requires_grad_(True)
    ref = torch.nn.CTCLoss(blank=0, zero_infinity=True)(lp_ref.transpose(0, 1), targets, il, tl)
    ref.backward()
    assert abs(ours.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert (lp.grad.double().cpu() - lp_ref.grad).abs().max().item() <= 2e-6 + 1e-4 * lp_ref.grad.abs().max().item()
