"""Per-kernel numerics on a real B200, through the C ABI (ctypes), against plain PyTorch fp32 references.

Tolerances: operands are bf16 (8 mantissa bits, eps 2^-8 = 3.9e-3), accumulation fp32.  A bf16-stored result
is therefore compared with  |got - ref| <= 1e-2 * max(|ref|) + small abs  (about 2.5 ulp of the largest value);
fp32-stored results with a tighter bound.  Integer / index work (SpecAugment) is compared bit-exactly.
"""
import math

import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from nn_conformer_for_speech_recognition_b200 import ops as _ops
    return _ops


def _close(got, ref, rel=1e-2, abs_=1e-3, name=""):
    got, ref = got.float(), ref.float()
    tol = rel * ref.abs().max().item() + abs_
    err = (got - ref).abs().max().item()
    assert math.isfinite(err) and err <= tol, f"{name}: max err {err:.3e} > tol {tol:.3e}"


def _bf(t):
    return t.to(torch.bfloat16)


# ----------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (256, 256, 256, 128), (300, 144, 144, 64),
                                      (1000, 576, 144, 128), (777, 1024, 256, 256), (129, 64, 1024, 64),
                                      (513, 768, 320, 0)])
def test_gemm_forward_kk(ops, M, N, K, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    out = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, block_n=bn)
    _close(out, ref, name="gemm kk bias")
    # SiLU epilogue + saved pre-activation
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=out2, block_n=bn)
    _close(out2, ref, name="gemm pre-act")
    _close(h, torch.nn.functional.silu(ref), name="gemm silu")
    # fp32 residual epilogue, in place: x = x + 0.5*(a w^T + b)
    x = torch.randn(M, N, device="cuda", generator=g)
    want = x + 0.5 * ref
    ops.gemm(a, w, M=M, N=N, K=K, bias=bias, resid=x, out=x, alpha=0.5, block_n=bn)
    _close(x, want, rel=5e-3, name="gemm resid")


@pytest.mark.parametrize("M,N,K", [(256, 256, 128), (300, 144, 576), (1000, 256, 1024)])
def test_gemm_dgrad_k_mn(ops, M, N, K):
    """dX[M,N] = dY[M,K] * W[K,N]  (W stored (out=K, in=N): B is MN-major)."""
    g = torch.Generator(device="cuda").manual_seed(7)
    dy = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K))
    ref = dy.float() @ w.float()
    out = ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True)
    _close(out, ref, name="gemm dgrad")
    # fused SiLU backward: out = (dy W) * silu'(u)
    u = _bf(torch.randn(M, N, device="cuda", generator=g))
    s = torch.sigmoid(u.float())
    out = ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u)
    _close(out, ref * (s * (1 + u.float() * (1 - s))), name="gemm silu bwd")


@pytest.mark.parametrize("M,N,K,split", [(256, 256, 512, 1), (144, 576, 1000, 4), (1024, 256, 3000, 16),
                                         (256, 1024, 23872, 37)])
def test_gemm_wgrad_mn_mn(ops, M, N, K, split):
    """dW[M,N] = dY[K,M]^T * X[K,N], reduction over tokens K, fp32 atomics across split-K."""
    g = torch.Generator(device="cuda").manual_seed(11)
    dy = _bf(torch.randn(K, M, device="cuda", generator=g))
    x = _bf(torch.randn(K, N, device="cuda", generator=g))
    ref = dy.float().T @ x.float()
    out = ops.gemm(dy, x, M=M, N=N, K=K, a_mn=True, b_mn=True, out_dtype=torch.float32, accumulate=True, split_k=split)
    _close(out, ref, rel=2e-3, name="gemm wgrad")


def test_gemm_dropout_is_regenerable(ops):
    M, N, K = 256, 256, 64
    a = _bf(torch.randn(M, K, device="cuda"))
    w = _bf(torch.randn(N, K, device="cuda"))
    o1 = ops.gemm(a, w, M=M, N=N, K=K, drop_p=0.5, drop_seed=123)
    o2 = ops.gemm(a, w, M=M, N=N, K=K, drop_p=0.5, drop_seed=123)
    o0 = ops.gemm(a, w, M=M, N=N, K=K)
    assert torch.equal(o1, o2)
    kept = o1 != 0
    frac = kept.float().mean().item()
    assert 0.45 < frac < 0.55
    _close(o1[kept], 2 * o0.float()[kept], name="dropout scale")
    # scale_cast regenerates the same mask for the backward pass
    ones = torch.ones(M, N, device="cuda")
    m = ops.scale_cast(ones, 1.0, 0.5, 123)
    assert torch.equal(m != 0, kept | (o0 == 0))


def test_gemm_saved_preactivation_encodes_dropout_for_backward(ops):
    """FFN first linear in training: h = drop(SiLU(u)), saved u carries B2C_DROPPED_PREACT where h was dropped, so the
    backward dgrad (SILU_BWD, alpha = 1/keep) yields exactly  (dY W) * mask/keep * SiLU'(u)  without regenerating the mask."""
    from nn_conformer_for_speech_recognition_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(9)
    M, N, K = 700, 512, 256
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w1 = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    u_true = a.float() @ w1.float().T + bias
    u_saved = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w1, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=u_saved, drop_p=0.5, drop_seed=11)
    keep = float(_lib.load().b2c_dropout_keep_scale(0.5))
    assert keep == 2.0
    dropped = u_saved.float() < -5e3
    assert 0.45 < dropped.float().mean().item() < 0.55
    assert torch.equal(h[dropped], torch.zeros_like(h[dropped]))
    _close(u_saved.float()[~dropped], u_true[~dropped], name="kept pre-activation")
    _close(h.float()[~dropped], keep * torch.nn.functional.silu(u_true)[~dropped], name="kept activation")
    dy = _bf(torch.randn(M, K, device="cuda", generator=g))
    w2 = _bf(torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K))        # (out=K, in=N): dgrad reads it MN-major
    du = ops.gemm(dy, w2, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u_saved, alpha=keep)
    sg = torch.sigmoid(u_true)
    want = (dy.float() @ w2.float()) * keep * (~dropped).float() * (sg * (1 + u_true * (1 - sg)))
    assert torch.equal(du[dropped], torch.zeros_like(du[dropped]))
    _close(du, want, rel=2e-2, name="du through saved pre-activation")


def test_gemm_rejects_bad_arguments(ops):
    a = _bf(torch.randn(128, 64, device="cuda"))
    with pytest.raises(ValueError):
        ops.gemm(a, a, M=128, N=100, K=64)          # N not a multiple of 8
    with pytest.raises(ValueError):
        ops.gemm(a.float(), a, M=128, N=128, K=64)  # wrong dtype
    with pytest.raises(ValueError):
        ops.gemm(a.cpu(), a, M=128, N=128, K=64)    # CPU tensor: no fallback


# ----------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("rows,d", [(1, 144), (37, 256), (1000, 512), (23872, 256), (65, 1024), (9, 16)])
def test_layernorm_fwd_bwd(ops, rows, d):
    g = torch.Generator(device="cuda").manual_seed(rows + d)
    x = (torch.randn(rows, d, device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, device="cuda", generator=g)).requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x, (d,), gamma, beta, 1e-5)
    y32, mean, rstd = ops.layernorm_fwd(x.detach(), gamma.detach(), beta.detach(), out_dtype=torch.float32, save_stats=True)
    _close(y32, ref, rel=1e-5, abs_=1e-5, name="ln fwd f32")
    y16 = ops.layernorm_fwd(x.detach(), gamma.detach(), beta.detach())
    _close(y16, ref, name="ln fwd bf16")
    dy = torch.randn(rows, d, device="cuda", generator=g)
    ref.backward(dy)
    for dyk in (dy, _bf(dy)):
        dx0 = torch.randn(rows, d, device="cuda", generator=g)
        dx = dx0.clone()
        dg = torch.zeros(d, device="cuda")
        db = torch.zeros(d, device="cuda")
        ops.layernorm_bwd(dyk, x.detach(), gamma.detach(), mean, rstd, dx, True, dg, db)
        rel = 1e-4 if dyk.dtype == torch.float32 else 1e-2
        _close(dx - dx0, x.grad, rel=rel, abs_=1e-4, name="ln dx")
        _close(dg, gamma.grad, rel=rel, abs_=1e-3, name="ln dgamma")
        _close(db, beta.grad, rel=rel, abs_=1e-3, name="ln dbeta")


@pytest.mark.parametrize("rows,d", [(1000, 256), (77, 144), (300, 512), (23872, 256), (5, 16)])
def test_layernorm_pair_kernels_equal_two_single_launches(ops, rows, d):
    """b2c_layernorm_fwd_pair / _bwd_pair (final LayerNorm of layer i + first LayerNorm of layer i+1 in one pass) against the
    two single-LayerNorm launches they replace: outputs, statistics, dx and the fused cast bit for bit; parameter
    gradients / column sums are atomics in another order (fp32 tolerance)."""
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + d)
    rnd = lambda *shape: torch.randn(*shape, device="cuda", generator=g)          # noqa: E731
    x = rnd(rows, d) * 1.7 + 0.4
    g1, b1, g2, b2 = 1 + 0.2 * rnd(d), 0.3 * rnd(d), 1 + 0.2 * rnd(d), 0.3 * rnd(d)
    # ---- forward
    y_ref, m1_ref, r1_ref = ops.layernorm_fwd(x, g1, b1, out_dtype=torch.float32, save_stats=True)
    a_ref, m2_ref, r2_ref = ops.layernorm_fwd(y_ref, g2, b2, save_stats=True)
    y, m1, r1, a, m2, r2 = ops.layernorm_fwd_pair(x, g1, b1, g2, b2, save_stats=True)
    for got, ref, name in ((y, y_ref, "y1"), (m1, m1_ref, "mean1"), (r1, r1_ref, "rstd1")):
        assert torch.equal(got, ref), name                # the first LayerNorm is the same instruction sequence
    for got, ref, name in ((m2, m2_ref, "mean2"), (r2, r2_ref, "rstd2")):
        _close(got, ref, rel=1e-6, abs_=1e-7, name=name)
    mism = (a.view(torch.int16) != a_ref.view(torch.int16)).float().mean().item()
    assert mism < 1e-3, f"a2: {mism:.2e} of the bf16 words differ"
    _close(a, a_ref, rel=1e-2, abs_=1e-2, name="a2")
    y_b, a_b = ops.layernorm_fwd_pair(x, g1, b1, g2, b2)                            # no statistics requested
    assert torch.equal(y_b, y) and torch.equal(a_b.view(torch.int16), a.view(torch.int16))
    # ---- backward
    da = _bf(rnd(rows, d))
    prev = rnd(rows, d)
    for drop_p in (0.0, 0.5):
        z = lambda: torch.zeros(d, device="cuda")                                   # noqa: E731
        dg2_r, db2_r, dg1_r, db1_r, cs_r = z(), z(), z(), z(), z()
        mid = prev.clone()
        ops.layernorm_bwd(da, y_ref, g2, m2_ref, r2_ref, mid, True, dg2_r, db2_r)
        dx_r = torch.empty_like(x)
        cast_r = ops.layernorm_bwd(mid, x, g1, m1_ref, r1_ref, dx_r, False, dg1_r, db1_r, cast=(0.5, drop_p, 91, cs_r))
        dg2, db2, dg1, db1, cs = z(), z(), z(), z(), z()
        dx = torch.empty_like(x)
        keep = prev.clone()
        cast = ops.layernorm_bwd_pair(da, y_ref, g2, m2_ref, r2_ref, dg2, db2, keep, x, g1, m1_ref, r1_ref, dg1, db1, dx,
                                      cast=(0.5, drop_p, 91, cs))
        assert torch.equal(keep, prev), "prev must only be read"
        _close(dx, dx_r, rel=1e-6, abs_=1e-6, name="pair dx")
        assert (cast.float() - cast_r.float()).abs().max().item() <= 1e-2 * cast_r.float().abs().max().item()
        assert ((cast.view(torch.int16) != cast_r.view(torch.int16)).float().mean().item()) < 1e-3   # same mask, same rounding
        for got, ref, name in ((dg2, dg2_r, "dgamma2"), (db2, db2_r, "dbeta2"), (dg1, dg1_r, "dgamma1"), (db1, db1_r, "dbeta1"), (cs, cs_r, "colsum")):
            _close(got, ref, rel=1e-4, abs_=1e-3, name=f"pair {name}")


# ----------------------------------------------------------------------------------------------- conv module chain
def test_layernorm_bwd_fused_cast_equals_separate_scale_cast(ops):
    """b2c_layernorm_bwd_cast = b2c_layernorm_bwd followed by b2c_scale_cast on the new dx, bit for bit (same dropout
    stream, same rounded column sums up to fp32 summation order)."""
    torch.manual_seed(4)
    rows, d = 1000, 256
    x = torch.randn(rows, d, device="cuda")
    gamma, beta = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    _, mean, rstd = ops.layernorm_fwd(x, gamma, beta, save_stats=True)
    dy = _bf(torch.randn(rows, d, device="cuda"))
    dx0 = torch.randn(rows, d, device="cuda")
    for drop_p in (0.0, 0.5):
        dxa, dxb = dx0.clone(), dx0.clone()
        dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        ops.layernorm_bwd(dy, x, gamma, mean, rstd, dxa, True, dg, db)
        cs_ref = torch.zeros(d, device="cuda")
        ref = ops.scale_cast(dxa, 0.5, drop_p, 77, colsum=cs_ref)
        dg2, db2, cs = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
        got = ops.layernorm_bwd(dy, x, gamma, mean, rstd, dxb, True, dg2, db2, cast=(0.5, drop_p, 77, cs))
        assert torch.equal(dxa, dxb)
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
        _close(cs, cs_ref, rel=1e-4, abs_=1e-3, name="fused cast colsum")
        _close(dg2, dg, rel=1e-4, abs_=1e-3, name="dgamma")


def _convmod_ref(u, wdw, bdw, bn_w, bn_b, mean, var, B, T, d, k):
    u = u.float().view(B, T, 2 * d)
    g = u[..., :d] * torch.sigmoid(u[..., d:])
    y = torch.nn.functional.conv1d(g.transpose(1, 2), wdw.view(d, 1, k), bdw, padding=(k - 1) // 2, groups=d)
    z = (y - mean[None, :, None]) / torch.sqrt(var[None, :, None] + 1e-5) * bn_w[None, :, None] + bn_b[None, :, None]
    return y.transpose(1, 2).reshape(B * T, d), torch.nn.functional.silu(z).transpose(1, 2).reshape(B * T, d)


@pytest.mark.parametrize("B,T,d,k", [(2, 50, 16, 7), (3, 373, 144, 31), (2, 129, 256, 31), (1, 128, 512, 33),
                                     (4, 1, 64, 3), (2, 300, 72, 5)])
def test_convmod_eval_and_train(ops, B, T, d, k):
    g = torch.Generator(device="cuda").manual_seed(B * T + d)
    u = _bf(torch.randn(B * T, 2 * d, device="cuda", generator=g))
    wdw = torch.randn(d, k, device="cuda", generator=g) / math.sqrt(k)
    bdw = 0.1 * torch.randn(d, device="cuda", generator=g)
    bn_w = 1 + 0.1 * torch.randn(d, device="cuda", generator=g)
    bn_b = 0.1 * torch.randn(d, device="cuda", generator=g)
    rmean = 0.1 * torch.randn(d, device="cuda", generator=g)
    rvar = 0.5 + torch.rand(d, device="cuda", generator=g)
    y_ref, c_ref = _convmod_ref(u, wdw, bdw, bn_w, bn_b, rmean, rvar, B, T, d, k)
    out = ops.convmod_eval(u, wdw, bdw, bn_w, bn_b, rmean, rvar, B, T)
    _close(out, c_ref, name="convmod eval")
    # train mode: batch statistics over all B*T positions, running stats updated with momentum 0.1
    y, stats = ops.convmod_dw_fwd(u, wdw, bdw, B, T)
    _close(y, y_ref, name="convmod dw")
    rm, rv = rmean.clone(), rvar.clone()
    mean, rstd = ops.convmod_bn_finalize(stats, B * T, rm, rv)
    # the statistics are those of the STORED (bf16) y, so BatchNorm and its backward see a self-consistent tensor
    mu_ref = y.float().mean(0)
    var_ref = y.float().var(0, unbiased=False)
    _close(mean, mu_ref, rel=1e-4, abs_=1e-5, name="bn mean")
    _close(rstd, 1 / torch.sqrt(var_ref + 1e-5), rel=1e-3, name="bn rstd")
    _close(mean, y_ref.mean(0), rel=5e-3, abs_=1e-3, name="bn mean vs fp32 chain")
    n = B * T
    _close(rm, 0.9 * rmean + 0.1 * mu_ref, rel=1e-4, abs_=1e-5, name="running mean")
    _close(rv, 0.9 * rvar + 0.1 * var_ref * n / max(n - 1, 1), rel=1e-3, name="running var")
    c = ops.convmod_bn_apply(y, mean, rstd, bn_w, bn_b)
    z = (y.float() - mu_ref) / torch.sqrt(var_ref + 1e-5) * bn_w + bn_b
    _close(c, torch.nn.functional.silu(z), name="bn apply")
    # finalize + apply as ONE launch (what the training forward uses): bit-identical outputs, statistics and running stats
    rm2, rv2 = rmean.clone(), rvar.clone()
    c2, mean2, rstd2 = ops.convmod_bn_finalize_apply(y, stats, B * T, bn_w, bn_b, rm2, rv2)
    assert torch.equal(mean2, mean) and torch.equal(rstd2, rstd), "fused finalize: statistics"
    assert torch.equal(rm2, rm) and torch.equal(rv2, rv), "fused finalize: running statistics"
    assert torch.equal(c2.view(torch.int16), c.view(torch.int16)), "fused finalize + apply: output"


# ----------------------------------------------------------------------------------------------- attention
def _attn_ref(qkv, lengths, B, T, H):
    q, k, v = (qkv.float().view(B, T, 3, H, 64)[:, :, i].transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / 8.0
    mask = torch.arange(T, device=qkv.device)[None, :] >= lengths[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ v).transpose(1, 2).reshape(B * T, H * 64), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,T,H,lens", [(1, 128, 1, [128]), (2, 100, 2, [100, 37]), (3, 373, 4, [373, 200, 1]),
                                        (2, 748, 2, [748, 513]), (2, 130, 1, [129, 130])])
def test_attention_fwd(ops, B, T, H, lens):
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = _bf(torch.randn(B * T, 3 * H * 64, device="cuda", generator=g))
    lengths = torch.tensor(lens, device="cuda", dtype=torch.int32)
    ref, lse_ref = _attn_ref(qkv, lengths, B, T, H)
    out, lse = ops.attention_fwd(qkv, lengths, B, T, H, 1 / 8.0, save_lse=True)
    _close(out, ref, rel=2e-2, name="attention out")
    _close(lse * math.log(2.0), lse_ref, rel=1e-3, abs_=1e-2, name="attention lse")


# rel_att: relative-positional scores (HF rel_shift).  Reference = tests/_rel_reference.py, pinned to the HF golden.
@pytest.mark.parametrize("B,T,H,lens", [(2, 10, 2, [10, 6]), (1, 128, 1, [128]), (2, 300, 2, [300, 131]),
                                        (2, 373, 4, [373, 200]), (1, 520, 1, [515])])
def test_attention_rel_fwd_bwd(ops, B, T, H, lens):
    from _rel_reference import rel_attention_core
    g = torch.Generator(device="cuda").manual_seed(T + 1)
    Hd, M, Rp = H * 64, B * T, (2 * T - 1 + 7) // 8 * 8
    qkv = _bf(0.7 * torch.randn(M, 4 * Hd, device="cuda", generator=g))
    pw = _bf(0.7 * torch.randn(Rp, Hd, device="cuda", generator=g))
    pw[2 * T - 1:] = 0
    lengths = torch.tensor(lens, device="cuda", dtype=torch.int32)
    dout = _bf(torch.randn(M, Hd, device="cuda", generator=g))
    qf = qkv.float().requires_grad_(True)
    pf = pw.float().requires_grad_(True)
    qu, k, v, qv = (qf.view(B, T, 4, H, 64)[:, :, i].transpose(1, 2) for i in range(4))
    o = rel_attention_core(qu, k, v, qv, pf[:2 * T - 1].view(2 * T - 1, H, 64).transpose(0, 1), lengths, 1 / 8.0)
    ref = o.transpose(1, 2).reshape(M, Hd)
    (ref * dout.float()).sum().backward()
    out, lse = ops.attention_rel_fwd(qkv, pw, lengths, B, T, H, 1 / 8.0, save_lse=True)
    _close(out, ref.detach(), rel=2e-2, name="rel attention out")
    dqkv, dpw = ops.attention_rel_bwd(qkv, pw, lengths, out, dout, lse, B, T, H, 1 / 8.0)
    valid = (torch.arange(T, device="cuda")[None, :] < lengths[:, None]).reshape(M, 1).float()
    for i, nm in enumerate(["d(q+u)", "dk", "dv", "d(q+v)"]):
        got, want = dqkv[:, i * Hd:(i + 1) * Hd].float(), qf.grad[:, i * Hd:(i + 1) * Hd]
        if i in (1, 2):       # keys/values past the utterance end receive no gradient
            assert float((got * (1 - valid)).abs().max()) == 0.0, nm
        _close(got, want, rel=3e-2, name="rel " + nm)
    _close(dpw, pf.grad, rel=3e-2, name="rel dPw")


def test_attention_dropout_mask_is_the_same_in_forward_and_backward(ops):
    """The backward kernel REGENERATES the probability-dropout mask from (seed, position).  With Q = K = 0 every
    probability is 1/T, so one-hot V rows read the forward mask off `out`, and one-hot dO rows read the backward mask
    off dV: the two must be identical (and about half the entries must survive at p = 0.5, scaled by 2)."""
    B, T, H = 2, 128, 2
    Hd = H * 64
    qkv = torch.zeros(B * T, 3 * Hd, device="cuda", dtype=torch.bfloat16)
    eye = torch.eye(64, device="cuda", dtype=torch.bfloat16)
    for b in range(B):
        for h in range(H):
            qkv[b * T:b * T + 64, 2 * Hd + h * 64:2 * Hd + (h + 1) * 64] = eye        # V rows of keys 0..63 are one-hot
    lengths = torch.full((B,), T, device="cuda", dtype=torch.int32)
    out, lse = ops.attention_fwd(qkv, lengths, B, T, H, 0.125, save_lse=True, drop_p=0.5, drop_seed=77)
    fwd = out.float().view(B, T, H, 64) * T                     # [b, i, h, j<64] = 2 * keep(i, j)
    assert set(fwd.unique().round().tolist()) <= {0.0, 2.0}
    assert 0.45 < (fwd != 0).float().mean().item() < 0.55
    dout = torch.zeros(B * T, Hd, device="cuda", dtype=torch.bfloat16)
    for b in range(B):
        for h in range(H):
            dout[b * T:b * T + 64, h * 64:(h + 1) * 64] = eye                        # dO rows of queries 0..63 are one-hot
    dqkv = ops.attention_bwd(qkv, lengths, out, dout, lse, B, T, H, 0.125, 0.5, 77)
    dv = dqkv[:, 2 * Hd:].float().view(B, T, H, 64) * T          # [b, j, h, i<64] = 2 * keep(i, j)
    bwd = dv.permute(0, 3, 2, 1)                                 # -> [b, i<64, h, j]
    assert torch.equal(bwd[:, :, :, :64] != 0, fwd[:, :64] != 0)
    assert 0.45 < (bwd != 0).float().mean().item() < 0.55
    out2 = ops.attention_fwd(qkv, lengths, B, T, H, 0.125, drop_p=0.5, drop_seed=78)
    assert not torch.equal(out2 != 0, out != 0)                  # another seed, another mask


def test_attention_zero_length_row_is_nan_like_torch(ops):
    qkv = _bf(torch.randn(2 * 64, 3 * 64, device="cuda"))
    lengths = torch.tensor([64, 0], device="cuda", dtype=torch.int32)
    out = ops.attention_fwd(qkv, lengths, 2, 64, 1, 0.125)
    assert torch.isfinite(out[:64]).all() and torch.isnan(out[64:].float()).all()


# ----------------------------------------------------------------------------------------------- ConvSubSampling
@pytest.mark.parametrize("B,F,T,C1,C2", [(2, 20, 30, 8, 16), (2, 80, 61, 24, 16), (3, 80, 300, 64, 128),
                                         (2, 40, 40, 512, 128)])
def test_convsub_compose_and_forward(ops, B, F, T, C1, C2):
    g = torch.Generator(device="cuda").manual_seed(C1)
    W1 = torch.randn(C1, 1, 7, 7, device="cuda", generator=g) / 7
    b1 = 0.1 * torch.randn(C1, device="cuda", generator=g)
    W2 = torch.randn(C2, C1, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C1)
    b2 = 0.1 * torch.randn(C2, device="cuda", generator=g)
    x = torch.rand(B, 1, F, T, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False       # keep the torch reference in true fp32
    ref = torch.nn.functional.conv2d(torch.nn.functional.conv2d(x, W1, b1, stride=2), W2, b2, stride=2)   # (B,C2,F',T')
    Wc, bc, Wc32 = ops.convsub_compose(W1, b1, W2, b2, 2, 128, want_f32=True)
    # composition identity in fp32
    refc = torch.nn.functional.conv2d(x, Wc32[:, :121].view(C2, 1, 11, 11), bc, stride=4)
    _close(refc, ref, rel=1e-3, abs_=1e-4, name="composition")
    out32 = ops.convsub_fwd(x[:, 0].contiguous(), Wc, bc, 11, 4, out_dtype=torch.float32)   # [B,T',F',C2]
    _close(out32.permute(0, 3, 2, 1), ref, name="convsub fwd f32")
    out16 = ops.convsub_fwd(x[:, 0].contiguous(), Wc, bc, 11, 4)
    _close(out16.permute(0, 3, 2, 1), ref, name="convsub fwd bf16")


def test_convsub_wgrad_and_compose_bwd(ops):
    B, F, T, C1, C2 = 2, 80, 200, 32, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    W1 = (torch.randn(C1, 1, 7, 7, device="cuda", generator=g) / 7).requires_grad_(True)
    b1 = (0.1 * torch.randn(C1, device="cuda", generator=g)).requires_grad_(True)
    W2 = (torch.randn(C2, C1, 3, 3, device="cuda", generator=g) / math.sqrt(9 * C1)).requires_grad_(True)
    b2 = (0.1 * torch.randn(C2, device="cuda", generator=g)).requires_grad_(True)
    x = torch.rand(B, 1, F, T, device="cuda", generator=g)
    torch.backends.cudnn.allow_tf32 = False
    ref = torch.nn.functional.conv2d(torch.nn.functional.conv2d(x, W1, b1, stride=2), W2, b2, stride=2)
    dref = torch.randn_like(ref)
    dref = _bf(dref).float()
    ref.backward(dref)
    dout = _bf(dref.permute(0, 3, 2, 1).contiguous())            # token-major [B,T',F',C2]
    Fo, To = ref.shape[2], ref.shape[3]
    dWc = torch.zeros(C2, 128, device="cuda")
    ops.convsub_wgrad(x[:, 0].contiguous(), dout.view(B * To, Fo * C2), dWc, 11, 4)
    dbc = torch.zeros(C2, device="cuda")
    ops.colsum(dout.view(-1, C2), dbc)
    _close(dbc, b2.grad, rel=2e-3, name="dbc")
    dW1, db1 = torch.zeros_like(W1), torch.zeros_like(b1)
    dW2, db2 = torch.zeros_like(W2), torch.zeros_like(b2)
    ops.convsub_compose_bwd(dWc, dbc, W1.detach(), b1.detach(), W2.detach(), 2, dW1, db1, dW2, db2)
    _close(dW1, W1.grad, name="dW1")
    _close(db1, b1.grad, rel=2e-3, name="db1")
    _close(dW2, W2.grad, name="dW2")
    _close(db2, b2.grad, rel=2e-3, name="db2")


# ----------------------------------------------------------------------------------------------- helpers
def test_cast_colsum_repack(ops):
    x = torch.randn(1001, 40, device="cuda")
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    a = _bf(torch.randn(5000, 144, device="cuda"))
    out = torch.zeros(144, device="cuda")
    ops.colsum(a, out)
    _close(out, a.float().sum(0), rel=1e-4, abs_=1e-2, name="colsum")
    src = _bf(torch.randn(10, 3 * 4 * 36, device="cuda"))
    dst = ops.repack_heads(src, 12, 36, 64)
    want = torch.zeros(10, 12, 64, device="cuda", dtype=torch.bfloat16)
    want[:, :, :36] = src.view(10, 12, 36)
    assert torch.equal(dst, want.view(10, 768))
    back = ops.repack_heads(dst, 12, 64, 36)
    assert torch.equal(back, src)


# ----------------------------------------------------------------------------------------------- fused Adafactor
def test_fused_adafactor_matches_hf():
    """Same update as transformers.Adafactor with the reference's arguments (runner.py:36), over several steps."""
    from transformers.optimization import Adafactor as HFAdafactor
    from nn_conformer_for_speech_recognition_b200.optim import Adafactor
    torch.manual_seed(0)
    shapes = [(256,), (1024, 256), (256, 1024), (512, 256, 1), (256, 1, 31), (512, 1, 7, 7), (128, 64, 3, 3), (144,),
              (256, 2304), (3, 5), (4, 8, 128), (2, 256, 128), (70, 384)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.1) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    o1 = Adafactor(ours, lr=2e-5, beta1=0.9)
    o2 = HFAdafactor(ref, lr=2e-5, beta1=0.9, scale_parameter=False, relative_step=False)
    for step in range(4):
        for a, b_ in zip(ours, ref):
            g = torch.randn_like(a) * (10.0 if step == 1 else 0.01)       # step 1 trips the RMS clip
            a.grad, b_.grad = g.clone(), g.clone()
        o1.step()
        o2.step()
        for i, (a, b_) in enumerate(zip(ours, ref)):
            err = (a.detach() - b_.detach()).abs().max().item()
            assert err <= 2e-7 + 1e-3 * 2e-5, f"step {step} tensor {shapes[i]}: |dp| {err:.3e}"
    for i, b_ in enumerate(ref):
        st = o2.state[b_]
        _close(o1.views[i]["exp_avg"], st["exp_avg"], rel=1e-3, abs_=1e-9, name="exp_avg")
        if b_.dim() >= 2:
            _close(o1.views[i]["row"].view(st["exp_avg_sq_row"].shape), st["exp_avg_sq_row"], rel=1e-3, abs_=1e-12, name="row")
            _close(o1.views[i]["col"].view(st["exp_avg_sq_col"].shape), st["exp_avg_sq_col"], rel=1e-3, abs_=1e-12, name="col")
        else:
            _close(o1.views[i]["row"], st["exp_avg_sq"], rel=1e-3, abs_=1e-12, name="sq")


@pytest.mark.parametrize("M,N,K", [(23872, 256, 1024), (23872, 1024, 256), (19100, 256, 256), (40000, 384, 128)])
def test_gemm_half_width_tail_tiles(ops, M, N, K):
    """Benchmark-size problems whose last wave of tiles is less than half full: that wave runs as half-width tiles
    (187 tiles on 148 SMs -> 148 full + 78 half).  Every epilogue family, forward and dgrad operand layouts."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    _close(ops.gemm(a, w, M=M, N=N, K=K, bias=bias), ref, name="tail bf16")
    u2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=u2)
    _close(u2, ref, name="tail pre-act")
    _close(h, torch.nn.functional.silu(ref), name="tail silu")
    x = torch.randn(M, N, device="cuda", generator=g)
    want = x + 0.5 * ref
    out = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, resid=x, out_dtype=torch.float32, alpha=0.5)
    _close(out, want, rel=5e-3, name="tail resid")
    wt = w.t().contiguous()                                   # [K, N]: the MN-major B of a dgrad
    _close(ops.gemm(a, wt, M=M, N=N, K=K, b_mn=True), a.float() @ wt.float(), name="tail dgrad")
    aux = _bf(torch.randn(M, N, device="cuda", generator=g))
    sg = torch.sigmoid(aux.float())
    cs = torch.zeros(N, device="cuda")
    du = ops.gemm(a, wt, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=aux, colsum=cs)
    want_du = (a.float() @ wt.float()) * (sg * (1 + aux.float() * (1 - sg)))
    _close(du, want_du, name="tail silu bwd")
    _close(cs, want_du.sum(0), rel=5e-3, abs_=0.5, name="tail colsum")


@pytest.mark.parametrize("M,N,K,bn", [(300, 144, 576, 256), (300, 144, 576, 128), (1000, 1000, 256, 256), (1000, 1000, 256, 128),
                                      (33, 40, 64, 128), (4100, 520, 128, 0), (23872, 1024, 256, 2256)])
def test_gemm_silu_backward_row_domain_ragged(ops, M, N, K, bn):
    """SiLU-backward dgrad whose saved pre-activation arrives as TMA boxes one chunk ahead: shapes where some of a
    warp's 32-column chunks lie past N (their loads / waits are skipped, buffer parities must stay in step), rows past
    M, the 1/keep factor and the fused bias gradient.  A strided aux (a column slice of a wider buffer) as well."""
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N)
    dy = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K))
    wide = _bf(torch.randn(M, N + 24, device="cuda", generator=g))
    u = wide[:, 8:8 + N]                                       # pitch N + 24 elements, 16-byte aligned start
    cs = torch.zeros(N, device="cuda")
    out = ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u, alpha=2.0, colsum=cs, block_n=bn)
    s = torch.sigmoid(u.float())
    want = 2.0 * (dy.float() @ w.float()) * (s * (1 + u.float() * (1 - s)))
    _close(out, want, name="silu bwd ragged")
    _close(cs, want.sum(0), rel=5e-3, abs_=0.5, name="silu bwd ragged colsum")


# ----------------------------------------------------------------------------------------------- GEMM, CTA pairs
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, -256), (512, 512, 256, -256), (1000, 1024, 256, -256),
                                      (777, 768, 320, -256), (300, 144, 144, -128), (23872, 256, 1024, -256),
                                      (2049, 384, 576, -128)])
def test_gemm_cta_pair_forward(ops, M, N, K, bn):
    """tcgen05.mma.cta_group::2 path (256 x |bn| tile per CTA pair), forward layout, all fused epilogues."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    _close(ops.gemm(a, w, M=M, N=N, K=K, bias=bias, block_n=bn), ref, name="pair bias")
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=out2, block_n=bn)
    _close(out2, ref, name="pair pre-act")
    _close(h, torch.nn.functional.silu(ref), name="pair silu")
    x = torch.randn(M, N, device="cuda", generator=g)
    want = x + 0.5 * ref
    ops.gemm(a, w, M=M, N=N, K=K, bias=bias, resid=x, out=x, alpha=0.5, block_n=bn)
    _close(x, want, rel=5e-3, name="pair resid")


# ----------------------------------------------------------------------------------------------- GEMM, multicast clusters
@pytest.mark.parametrize("M,N,K,bn", [(256, 256, 64, 2256), (1000, 1024, 256, 2256), (777, 768, 320, 4256),
                                      (23872, 256, 1024, 4256), (300, 144, 144, 2128), (2049, 384, 576, 4128),
                                      (23872, 1024, 256, 4256), (130, 256, 256, 4256)])
def test_gemm_multicast_cluster_forward(ops, M, N, K, bn):
    """block_n = 1000*CL + BN: CL CTAs on vertically adjacent tiles share one TMA-multicast B tile; ragged M leaves
    whole CTAs of a cluster without rows (they still take part in the multicast)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    _close(ops.gemm(a, w, M=M, N=N, K=K, bias=bias, block_n=bn), ref, name="cluster bias")
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=out2, block_n=bn)
    _close(out2, ref, name="cluster pre-act")
    _close(h, torch.nn.functional.silu(ref), name="cluster silu")
    x = torch.randn(M, N, device="cuda", generator=g)
    want = x + 0.5 * ref
    ops.gemm(a, w, M=M, N=N, K=K, bias=bias, resid=x, out=x, alpha=0.5, block_n=bn)
    _close(x, want, rel=5e-3, name="cluster resid")


# ----------------------------------------------------------------------------------------------- GEMM, weight-stationary
@pytest.mark.parametrize("M,N,K,bn", [(512, 256, 64, 9256), (1000, 1024, 256, 9256), (777, 768, 144, 9256), (777, 384, 256, 9128),
                                      (23872, 1024, 256, 9128), (23872, 768, 256, 9256), (2049, 520, 200, 9128), (130, 256, 256, 9256)])
def test_gemm_weight_stationary(ops, M, N, K, bn):
    """block_n = 9000 + BN (K <= 256): every CTA parks one n-block of the weights in shared memory and streams A only.
    Ragged M / N, a remainder round issued as half-width tiles, every row-domain epilogue incl. the one-bit dropout
    masks with the saved pre-activation, and the MN-major (dgrad) weight layout with the SiLU-backward epilogue."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    _close(ops.gemm(a, w, M=M, N=N, K=K, bias=bias, block_n=bn), ref, name="stationary bias")
    _close(ops.gemm(a, w, M=M, N=N, K=K, block_n=bn), ref - bias, name="stationary no bias")
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    h = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=out2, block_n=bn)
    _close(out2, ref, name="stationary pre-act")
    _close(h, torch.nn.functional.silu(ref), name="stationary silu")
    _close(ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, block_n=bn), torch.nn.functional.silu(ref), name="stationary silu eval")
    for p in (0.5, 0.25):
        hd = ops.gemm(a, w, M=M, N=N, K=K, bias=bias, act=ops.ACT_SILU, out2=out2, drop_p=p, drop_seed=77, block_n=bn)
        dropped = out2.float() < -5e3
        assert abs(dropped.float().mean().item() - p) < 0.03
        assert torch.equal(hd[dropped], torch.zeros_like(hd[dropped]))
        _close(out2.float()[~dropped], ref[~dropped], name="stationary kept pre-act")
        _close(hd.float()[~dropped], torch.nn.functional.silu(ref)[~dropped] / (1 - p), rel=2e-2, name="stationary kept act")
    wt = _bf(torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K))       # [K, N] storage: dgrad layout
    _close(ops.gemm(a, wt, M=M, N=N, K=K, b_mn=True, block_n=bn), a.float() @ wt.float(), name="stationary dgrad")
    u = _bf(torch.randn(M, N, device="cuda", generator=g))
    cs = torch.zeros(N, device="cuda")
    du = ops.gemm(a, wt, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u, alpha=2.0, colsum=cs, block_n=bn)
    sg = torch.sigmoid(u.float())
    want = 2.0 * (a.float() @ wt.float()) * (sg * (1 + u.float() * (1 - sg)))
    _close(du, want, name="stationary silu bwd")
    _close(cs, want.sum(0), rel=5e-3, abs_=0.5, name="stationary silu bwd colsum")


@pytest.mark.parametrize("bn", [2128, 2256, 4256])
def test_gemm_multicast_cluster_dgrad(ops, bn):
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 3000, 512, 256
    dy = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K))
    _close(ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, block_n=bn), dy.float() @ w.float(), name="cluster dgrad")
    u = _bf(torch.randn(M, N, device="cuda", generator=g))
    s = torch.sigmoid(u.float())
    out = ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u, block_n=bn)
    _close(out, (dy.float() @ w.float()) * (s * (1 + u.float() * (1 - s))), name="cluster silu bwd")


@pytest.mark.parametrize("bn", [-256, -128])
def test_gemm_cta_pair_dgrad_wgrad(ops, bn):
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 1000, 512, 256
    dy = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K))
    _close(ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, block_n=bn), dy.float() @ w.float(), name="pair dgrad")
    u = _bf(torch.randn(M, N, device="cuda", generator=g))
    s = torch.sigmoid(u.float())
    out = ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u, block_n=bn)
    _close(out, (dy.float() @ w.float()) * (s * (1 + u.float() * (1 - s))), name="pair silu bwd")
    Mo, No, Kt = 256, 1024, 5000
    dyt = _bf(torch.randn(Kt, Mo, device="cuda", generator=g))
    xt = _bf(torch.randn(Kt, No, device="cuda", generator=g))
    dw = ops.gemm(dyt, xt, M=Mo, N=No, K=Kt, a_mn=True, b_mn=True, out_dtype=torch.float32, accumulate=True, split_k=7,
                  block_n=bn)
    _close(dw, dyt.float().T @ xt.float(), rel=2e-3, name="pair wgrad")


def test_fused_column_sums(ops):
    """bias gradients are produced by the kernels that write the tensor (scale_cast, GEMM epilogue)."""
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(5000, 256, device="cuda", generator=g)
    cs = torch.zeros(256, device="cuda")
    y = ops.scale_cast(x, 0.5, 0.5, 77, colsum=cs)
    _close(cs, y.float().sum(0), rel=1e-4, abs_=1e-2, name="scale_cast colsum")
    assert torch.equal(y, ops.scale_cast(x, 0.5, 0.5, 77))
    x2 = torch.randn(333, 144, device="cuda", generator=g)
    cs2 = torch.zeros(144, device="cuda")
    y2 = ops.scale_cast(x2, 1.0, colsum=cs2)
    assert torch.equal(y2, x2.to(torch.bfloat16))
    _close(cs2, y2.float().sum(0), rel=1e-4, abs_=1e-2, name="scale_cast colsum 144")
    M, N, K = 3000, 1024, 256
    dy = _bf(torch.randn(M, K, device="cuda", generator=g))
    w = _bf(torch.randn(K, N, device="cuda", generator=g) / 16)
    u = _bf(torch.randn(M, N, device="cuda", generator=g))
    sg = torch.sigmoid(u.float())
    exact = ((dy.float() @ w.float()) * (sg * (1 + u.float() * (1 - sg)))).sum(0)      # the GEMM sums its fp32 values
    for bn in (128, -256, 64):
        cs3 = torch.zeros(N, device="cuda")
        out = ops.gemm(dy, w, M=M, N=N, K=K, b_mn=True, act=ops.ACT_SILU_BWD, aux=u, colsum=cs3, block_n=bn)
        _close(cs3, exact, rel=3e-3, abs_=5e-2, name=f"gemm colsum bn{bn}")
        _close(cs3, out.float().sum(0), rel=2e-2, abs_=5e-2, name=f"gemm colsum vs rounded outputs bn{bn}")


@pytest.mark.parametrize("M,d,f", [(300, 256, 1024), (1000, 144, 576), (129, 128, 256), (64, 16, 32), (5000, 256, 1024),
                                   (23872, 256, 1024), (640, 256, 136)])
def test_ffn_fused_eval(ops, M, d, f):
    """out = x + 0.5 * (W2 SiLU(W1 LN(x) + b1) + b2) in ONE kernel (torchaudio conformer.py:102-109 + :185-187, eval) against
    fp64 torch on the bf16-rounded weights, and against the un-fused LN / GEMM / GEMM path it replaces."""
    torch.manual_seed(M + d)
    x = torch.randn(M, d, device="cuda") * 1.5 + 0.3
    ln_w, ln_b = torch.rand(d, device="cuda") + 0.5, torch.randn(d, device="cuda") * 0.1
    w1, b1 = _bf(torch.randn(f, d, device="cuda") / math.sqrt(d)), torch.randn(f, device="cuda") * 0.1
    w2, b2 = _bf(torch.randn(d, f, device="cuda") / math.sqrt(f)), torch.randn(d, device="cuda") * 0.1
    out = ops.ffn_fused_eval(x, ln_w, ln_b, w1, b1, w2, b2)
    xd = x.double()
    a = torch.nn.functional.layer_norm(xd, (d,), ln_w.double(), ln_b.double(), 1e-5)
    ref = xd + 0.5 * (torch.nn.functional.silu(a @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double())
    assert torch.isfinite(out).all()
    err = (out.double() - ref).abs().max().item()
    assert err <= 1.5e-2 * (ref - xd).abs().max().item() + 1e-3, f"max err {err:.3e}"
    # un-fused path (what training and d > 256 use)
    a_bf = ops.layernorm_fwd(x, ln_w, ln_b)
    h = ops.gemm(a_bf, w1, M=M, N=f, K=d, bias=b1, act=ops.ACT_SILU)
    un = ops.gemm(h, w2, M=M, N=d, K=f, bias=b2, resid=x, out_dtype=torch.float32, alpha=0.5)
    assert (out - un).abs().max().item() <= 1e-2 * (ref - xd).abs().max().item() + 1e-3
    # in place
    x2 = x.clone()
    ops.ffn_fused_eval(x2, ln_w, ln_b, w1, b1, w2, b2, out=x2)
    assert torch.equal(x2, out)
