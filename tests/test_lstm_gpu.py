"""b2c.LSTM (the reference's bi-LSTM decoder, /root/reference/lib/standard/asrnn.py:38,252) against torch.nn.LSTM in fp32:
outputs, final states, and every gradient.  GEMM operands are bf16 (fp32 accumulation / cell state), so the tolerance is the
bf16 one of the other module tests: |got - ref| <= 2e-2 * max|ref| for activations, cosine >= 0.999 for gradients."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _close(got, ref, rel=2e-2, abs_=2e-3, name=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    tol = rel * ref.abs().max().item() + abs_
    err = (got - ref).abs().max().item()
    assert err <= tol, f"{name}: max err {err:.3e} > tol {tol:.3e}"


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("L,N,In,H,bidir,hx", [(5, 7, 16, 8, False, False), (9, 33, 64, 32, True, False), (12, 130, 256, 64, True, True),
                                                (64, 40, 256, 512, True, False)])
def test_lstm_forward_backward_vs_torch(L, N, In, H, bidir, hx):
    import nn_conformer_for_speech_recognition_b200 as b2c
    torch.manual_seed(L * 1000 + N)
    ref = torch.nn.LSTM(In, H, bidirectional=bidir)
    ours = b2c.lstm_from_reference(ref).cuda()
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    dirs = 2 if bidir else 1
    x = torch.randn(L, N, In)
    h0 = torch.randn(dirs, N, H) * 0.5 if hx else None
    c0 = torch.randn(dirs, N, H) * 0.5 if hx else None
    go = torch.randn(L, N, dirs * H)
    gh, gc = torch.randn(dirs, N, H), torch.randn(dirs, N, H)

    def run(m, dev):
        xx = x.detach().clone().to(dev).requires_grad_(True)
        st = None
        if hx:
            st = (h0.detach().clone().to(dev).requires_grad_(True), c0.detach().clone().to(dev).requires_grad_(True))
        out, (hn, cn) = m(xx, st)
        loss = (out * go.to(dev)).sum() + (hn * gh.to(dev)).sum() + (cn * gc.to(dev)).sum()
        m.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        grads["x"] = xx.grad
        if hx:
            grads["h0"], grads["c0"] = st[0].grad, st[1].grad
        return out.detach(), hn.detach(), cn.detach(), grads

    o_r, h_r, c_r, g_r = run(ref, "cpu")
    o_o, h_o, c_o, g_o = run(ours, "cuda")
    _close(o_o, o_r, name="output")
    _close(h_o, h_r, name="h_n")
    _close(c_o, c_r, name="c_n")
    for k, g in g_r.items():
        c = _cos(g_o[k], g)
        assert c >= 0.999, f"grad {k}: cosine {c:.5f}"
        _close(g_o[k], g, rel=3e-2, abs_=1e-2, name=f"grad {k}")


def test_lstm_batch_first_eval_and_reference_shape():
    """the reference's call: x = self.lstm(x)[0] on [batch, max_len, 256] with hidden 512 per direction (asrnn.py:38,252)"""
    import nn_conformer_for_speech_recognition_b200 as b2c
    torch.manual_seed(3)
    ref = torch.nn.LSTM(256, 512, bidirectional=True)
    ours = b2c.lstm_from_reference(ref).cuda().eval()
    x = torch.randn(8, 373, 256)
    with torch.no_grad():
        want = ref(x)[0]
        got = ours(x.cuda())[0]
    assert got.shape == want.shape == (8, 373, 1024)
    _close(got, want, name="reference-shape output")
    bf = b2c.LSTM(64, 32, batch_first=True).cuda()
    rf = torch.nn.LSTM(64, 32, batch_first=True)
    rf.load_state_dict({k: v.cpu() for k, v in bf.state_dict().items()})
    xb = torch.randn(6, 11, 64)
    with torch.no_grad():
        _close(bf(xb.cuda())[0], rf(xb)[0], name="batch_first output")


def test_lstm_rejects_what_the_reference_does_not_build():
    import nn_conformer_for_speech_recognition_b200 as b2c
    with pytest.raises(NotImplementedError):
        b2c.LSTM(64, 32, num_layers=2)
    with pytest.raises(NotImplementedError):
        b2c.LSTM(64, 32, proj_size=16)
