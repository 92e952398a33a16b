"""Thin tensor-level wrappers over the C ABI (one call = one kernel launch on the current torch stream).

Nothing here computes with torch ops: tensors are only allocated (torch.empty / zeros) and their device pointers
handed to libb200conformer.so.  Shapes follow include/b200conformer.h.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_SILU, ACT_SILU_BWD, BF16, F32, Epilogue, call, ptr, stream_ptr

__all__ = ["gemm", "layernorm_fwd", "layernorm_bwd", "convmod_eval", "convmod_dw_fwd", "convmod_bn_finalize",
           "convmod_bn_apply", "attention_fwd", "attention_rel_fwd", "attention_rel_bwd", "convsub_compose", "convsub_fwd", "convsub_wgrad",
           "convsub_compose_bwd", "specaug_apply", "cast_bf16", "scale_cast", "colsum", "repack_heads",
           "ACT_NONE", "ACT_SILU", "ACT_SILU_BWD"]


def _check(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, *, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, aux: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, out2: Optional[torch.Tensor] = None,
         out_dtype: torch.dtype = torch.bfloat16, alpha: float = 1.0, drop_p: float = 0.0, drop_seed: int = 0,
         accumulate: bool = False, split_k: int = 1, block_n: int = 0,
         colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """D[M,N] = epilogue(A * B^T) -- see b2c_gemm_bf16.  a/b are 2-D bf16 tensors whose last dim is contiguous;
    with a_mn/b_mn the tensor is the [K, rows] storage."""
    _lib.require_device()
    _check(a, torch.bfloat16, "a")
    _check(b, torch.bfloat16, "b")
    if a.stride(-1) != 1 or b.stride(-1) != 1:
        raise ValueError("gemm operands must have a contiguous last dimension")
    if out is None:
        out = (torch.zeros if accumulate else torch.empty)((M, N), device=a.device, dtype=out_dtype)
    ep = Epilogue()
    ep.bias = ptr(bias)
    ep.aux = ptr(aux)
    ep.resid = ptr(resid)
    ep.out = ptr(out)
    ep.out2 = ptr(out2)
    ep.ld_out = out.stride(0)
    ep.ld_out2 = out2.stride(0) if out2 is not None else 0
    ep.ld_aux = aux.stride(0) if aux is not None else 0
    ep.alpha = alpha
    ep.drop_p = drop_p
    ep.drop_seed = drop_seed
    ep.act = act
    ep.out_dtype = BF16 if out.dtype == torch.bfloat16 else F32
    ep.accumulate = int(accumulate)
    ep.colsum = ptr(colsum)
    # algorithmic HBM bytes of the launch: both operands once, every epilogue operand once, the output(s) once
    osz = 2 if out.dtype == torch.bfloat16 else 4
    nbytes = 2.0 * (M * K + N * K) + M * N * (osz * (2 if accumulate else 1) + (2 if out2 is not None else 0)
                                             + (2 if aux is not None else 0) + (4 if resid is not None else 0))
    call("b2c_gemm_bf16", ptr(a), a.stride(0), int(a_mn), ptr(b), b.stride(0), int(b_mn), M, N, K, split_k, block_n,
         ctypes.byref(ep), stream_ptr(), flops=2.0 * M * N * K, nbytes=nbytes)
    return out


def ffn_fused_eval(x, ln_w, ln_b, w1_bf16, b1, w2_bf16, b2, alpha: float = 0.5, eps: float = 1e-5, out=None):
    """out = x + alpha * (W2 SiLU(W1 LN(x) + b1) + b2) in one kernel (inference); x fp32 [M, d], may be updated in place."""
    _lib.require_device()
    _check(x, torch.float32, "x")
    M, d = x.shape
    f = w1_bf16.shape[0]
    if out is None:
        out = torch.empty_like(x)
    call("b2c_ffn_fused_eval", ptr(x), ptr(out), ptr(ln_w), ptr(ln_b), ptr(w1_bf16), ptr(b1), ptr(w2_bf16), ptr(b2), M, d, f,
         alpha, eps, stream_ptr(), flops=4.0 * M * d * f)
    return out


def ffn_fused_ok(d: int, f: int) -> bool:
    return d % 16 == 0 and d <= 256 and f % 8 == 0


def layernorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, out_dtype=torch.bfloat16,
                  save_stats: bool = False, eps: float = 1e-5, out: Optional[torch.Tensor] = None):
    _lib.require_device()
    _check(x, torch.float32, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    call("b2c_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(out), BF16 if out.dtype == torch.bfloat16 else F32,
         ptr(mean), ptr(rstd), rows, d, eps, stream_ptr())
    return (out, mean, rstd) if save_stats else out


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, accumulate: bool, dgamma, dbeta, cast=None):
    """cast = (alpha, drop_p, drop_seed, colsum): also return bf16(alpha * dropout(dx_new)) (what `scale_cast(dx, ...)`
    would compute in a second pass) with its column sums accumulated into `colsum`."""
    d = x.shape[-1]
    rows = x.numel() // d
    if cast is None:
        call("b2c_layernorm_bwd", ptr(dy), BF16 if dy.dtype == torch.bfloat16 else F32, ptr(x), ptr(gamma), ptr(mean),
             ptr(rstd), ptr(dx), int(accumulate), ptr(dgamma), ptr(dbeta), rows, d, stream_ptr())
        return dx
    alpha, drop_p, drop_seed, colsum = cast
    out = torch.empty((rows, d), device=dx.device, dtype=torch.bfloat16)
    call("b2c_layernorm_bwd_cast", ptr(dy), BF16 if dy.dtype == torch.bfloat16 else F32, ptr(x), ptr(gamma), ptr(mean),
         ptr(rstd), ptr(dx), int(accumulate), ptr(dgamma), ptr(dbeta), rows, d, ptr(out), alpha, drop_p, drop_seed,
         ptr(colsum), stream_ptr())
    return out


def layernorm_fwd_pair(x, gamma1, beta1, gamma2, beta2, *, save_stats: bool = False, eps: float = 1e-5, out1=None, out2=None):
    """y1 = LN1(x) (fp32), a2 = LN2(y1) (bf16) in one pass -- see b2c_layernorm_fwd_pair.
    Returns (y1, a2) or (y1, mean1, rstd1, a2, mean2, rstd2)."""
    _lib.require_device()
    _check(x, torch.float32, "x")
    d = x.shape[-1]
    rows = x.numel() // d
    y1 = out1 if out1 is not None else torch.empty(x.shape, device=x.device, dtype=torch.float32)
    a2 = out2 if out2 is not None else torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    m1 = r1 = m2 = r2 = None
    if save_stats:
        m1, r1, m2, r2 = (torch.empty(rows, device=x.device, dtype=torch.float32) for _ in range(4))
    call("b2c_layernorm_fwd_pair", ptr(x), ptr(gamma1), ptr(beta1), ptr(y1), ptr(m1), ptr(r1), ptr(gamma2), ptr(beta2), ptr(a2),
         ptr(m2), ptr(r2), rows, d, eps, stream_ptr())
    return (y1, m1, r1, a2, m2, r2) if save_stats else (y1, a2)


def layernorm_bwd_pair(da2, y1, gamma2, mean2, rstd2, dgamma2, dbeta2, prev, x, gamma1, mean1, rstd1, dgamma1, dbeta1, dx, cast):
    """Backward of `layernorm_fwd_pair`: dx = LN1'(prev + LN2'(da2)), returns bf16(alpha * dropout(dx)) -- see
    b2c_layernorm_bwd_pair.  cast = (alpha, drop_p, drop_seed, colsum)."""
    _check(da2, torch.bfloat16, "da2")
    d = x.shape[-1]
    rows = x.numel() // d
    alpha, drop_p, drop_seed, colsum = cast
    out = torch.empty((rows, d), device=dx.device, dtype=torch.bfloat16)
    call("b2c_layernorm_bwd_pair", ptr(da2), ptr(y1), ptr(gamma2), ptr(mean2), ptr(rstd2), ptr(dgamma2), ptr(dbeta2), ptr(prev),
         ptr(x), ptr(gamma1), ptr(mean1), ptr(rstd1), ptr(dgamma1), ptr(dbeta1), ptr(dx), rows, d, ptr(out), alpha, drop_p,
         drop_seed, ptr(colsum), stream_ptr())
    return out


def convmod_eval(u, wdw, bdw, bn_w, bn_b, bn_mean, bn_var, B: int, T: int, eps: float = 1e-5):
    _lib.require_device()
    _check(u, torch.bfloat16, "u")
    d, k = wdw.shape[0], wdw.shape[-1]
    out = torch.empty((B * T, d), device=u.device, dtype=torch.bfloat16)
    call("b2c_convmod_eval_fwd", ptr(u), ptr(wdw), ptr(bdw), ptr(bn_w), ptr(bn_b), ptr(bn_mean), ptr(bn_var), eps,
         ptr(out), B, T, d, k, stream_ptr(), nbytes=3.0 * B * T * d * 2)
    return out


def convmod_dw_fwd(u, wdw, bdw, B: int, T: int, stats=None):
    """stats: optional ZEROED fp64 [2d] slice (the caller zeroes one buffer for all layers with a single fill)."""
    d, k = wdw.shape[0], wdw.shape[-1]
    y = torch.empty((B * T, d), device=u.device, dtype=torch.bfloat16)
    if stats is None:
        stats = torch.zeros(2 * d, device=u.device, dtype=torch.float64)
    call("b2c_convmod_dw_fwd", ptr(u), ptr(wdw), ptr(bdw), ptr(y), ptr(stats), B, T, d, k, stream_ptr(),
         nbytes=3.0 * B * T * d * 2)
    return y, stats


def convmod_bn_finalize(stats, count: int, running_mean, running_var, eps: float = 1e-5, momentum: float = 0.1):
    d = stats.numel() // 2
    mean = torch.empty(d, device=stats.device, dtype=torch.float32)
    rstd = torch.empty(d, device=stats.device, dtype=torch.float32)
    call("b2c_convmod_bn_finalize", ptr(stats), count, eps, momentum, ptr(mean), ptr(rstd), ptr(running_mean),
         ptr(running_var), d, stream_ptr())
    return mean, rstd


def convmod_bn_finalize_apply(y, stats, count: int, bn_w, bn_b, running_mean, running_var, eps: float = 1e-5, momentum: float = 0.1):
    """`convmod_bn_finalize` + `convmod_bn_apply` in one launch; returns (out, mean, rstd)."""
    d = y.shape[1]
    mean = torch.empty(d, device=y.device, dtype=torch.float32)
    rstd = torch.empty(d, device=y.device, dtype=torch.float32)
    out = torch.empty_like(y)
    call("b2c_convmod_bn_finalize_apply", ptr(y), ptr(stats), count, eps, momentum, ptr(bn_w), ptr(bn_b), ptr(out), ptr(mean),
         ptr(rstd), ptr(running_mean), ptr(running_var), y.shape[0], d, stream_ptr(), nbytes=2.0 * y.shape[0] * d * 2)
    return out, mean, rstd


def convmod_bn_apply(y, mean, rstd, bn_w, bn_b):
    out = torch.empty_like(y)
    call("b2c_convmod_bn_apply", ptr(y), ptr(mean), ptr(rstd), ptr(bn_w), ptr(bn_b), ptr(out), y.shape[0], y.shape[1],
         stream_ptr())
    return out


def attention_fwd(qkv: torch.Tensor, lengths_i32: torch.Tensor, B: int, T: int, H: int, scale: float, *,
                  save_lse: bool = False, drop_p: float = 0.0, drop_seed: int = 0):
    """qkv bf16 [B*T, 3*H*64] (heads padded to 64) -> out bf16 [B*T, H*64]."""
    _lib.require_device()
    _check(qkv, torch.bfloat16, "qkv")
    _check(lengths_i32, torch.int32, "lengths")
    out = torch.empty((B * T, H * 64), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32) if save_lse else None
    call("b2c_attention_fwd", ptr(qkv), qkv.stride(0), H * 64, 2 * H * 64, ptr(lengths_i32), ptr(out), out.stride(0),
         ptr(lse), B, T, H, scale, drop_p, drop_seed, stream_ptr(), flops=4.0 * B * H * T * T * 64)
    return (out, lse) if save_lse else out


def attention_rel_fwd(qkv: torch.Tensor, pw: torch.Tensor, lengths_i32: torch.Tensor, B: int, T: int, H: int,
                      scale: float, *, save_lse: bool = False, drop_p: float = 0.0, drop_seed: int = 0):
    """Relative-positional attention (HF `_apply_relative_embeddings`).  qkv bf16 [B*T, 4*H*64] = [q+u | k | v | q+v];
    pw bf16 [>= 2T-1, H*64] projected relative embeddings, row r = relative position T-1-r."""
    _lib.require_device()
    _check(qkv, torch.bfloat16, "qkv")
    _check(pw, torch.bfloat16, "pw")
    _check(lengths_i32, torch.int32, "lengths")
    out = torch.empty((B * T, H * 64), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32) if save_lse else None
    call("b2c_attention_rel_fwd", ptr(qkv), qkv.stride(0), H * 64, 2 * H * 64, 3 * H * 64, ptr(pw), pw.stride(0),
         pw.shape[0], ptr(lengths_i32), ptr(out), out.stride(0), ptr(lse), B, T, H, scale, drop_p, drop_seed, stream_ptr(),
         flops=4.0 * B * H * T * T * 64 + 2.0 * B * H * T * (2 * T) * 64)
    return (out, lse) if save_lse else out


def attention_rel_bwd(qkv, pw, lengths_i32, att, datt, lse, B: int, T: int, H: int, scale: float, drop_p: float = 0.0,
                      drop_seed: int = 0):
    """-> (dqkv bf16 [B*T, 4*H*64] = [d(q+u) | dk | dv | d(q+v)], dpw fp32 [Rp, H*64]).  BD is re-materialised with the
    GEMM kernel (one launch per head), the attention kernel scatters dS to the shifted index, and the two products that
    consume it are GEMMs again (dgrad / split-K wgrad form)."""
    Hd = H * 64
    M = B * T
    Rp = pw.shape[0]
    dev = qkv.device
    bd = torch.empty((M, H, Rp), device=dev, dtype=torch.float32)
    for h in range(H):
        gemm(qkv[:, 3 * Hd + h * 64:3 * Hd + (h + 1) * 64], pw[:, h * 64:(h + 1) * 64], M=M, N=Rp, K=64, out=bd[:, h, :])
    dbd = torch.zeros((M, H, Rp), device=dev, dtype=torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, T), device=dev, dtype=torch.float32)
    dq32 = torch.empty((M, Hd), device=dev, dtype=torch.float32)          # zeroed by the library, accumulated by TMA reduce-adds
    call("b2c_attention_rel_bwd", ptr(qkv), qkv.stride(0), Hd, 2 * Hd, ptr(lengths_i32), ptr(att), ptr(datt),
         att.stride(0), ptr(lse), ptr(delta), ptr(dq32), ptr(dqkv), ptr(bd), ptr(dbd), Rp, B, T, H, scale, drop_p,
         drop_seed, stream_ptr(), flops=14.0 * B * H * T * T * 64)
    dpw = torch.zeros((Rp, Hd), device=dev, dtype=torch.float32)
    tiles = (Rp + 127) // 128
    kb = (M + 63) // 64
    split = max(1, min(kb // 4 if kb >= 4 else 1, (296 + tiles - 1) // tiles))
    for h in range(H):
        dbd_h = dbd[:, h, :]
        qv_h = qkv[:, 3 * Hd + h * 64:3 * Hd + (h + 1) * 64]
        # d(q+v)[M,64] = dBD_h[M,Rp] Pw_h[Rp,64]
        gemm(dbd_h, pw[:, h * 64:(h + 1) * 64], M=M, N=64, K=Rp, b_mn=True, out=dqkv[:, 3 * Hd + h * 64:3 * Hd + (h + 1) * 64])
        # dPw_h[Rp,64] = dBD_h^T (q+v)_h
        gemm(dbd_h, qv_h, M=Rp, N=64, K=M, a_mn=True, b_mn=True, out=dpw[:, h * 64:(h + 1) * 64], accumulate=True,
             split_k=split)
    return dqkv, dpw


def convsub_compose(W1, b1, W2, b2, s1: int, KP: int, want_f32: bool = False, out=None):
    """out = (Wc bf16 [>=C2, KP], bc fp32 [>=C2]): write into these (static) buffers instead of allocating."""
    _lib.require_device()
    C1, _, k1, _ = W1.shape
    C2, _, k2, _ = W2.shape
    if out is not None:
        Wc, bc = out
    else:
        Wc = torch.empty((C2, KP), device=W1.device, dtype=torch.bfloat16)
        bc = torch.empty(C2, device=W1.device, dtype=torch.float32)
    Wc32 = torch.empty((C2, KP), device=W1.device, dtype=torch.float32) if want_f32 else None
    call("b2c_convsub_compose", ptr(W1), ptr(b1), ptr(W2), ptr(b2), C1, C2, k1, s1, k2, ptr(Wc), ptr(Wc32), ptr(bc), KP,
         stream_ptr())
    return (Wc, bc, Wc32) if want_f32 else (Wc, bc)


def convsub_fwd(x, Wc, bc, kc: int, stride: int, out_dtype=torch.bfloat16):
    """x fp32 [B,F,T] -> [B, T', F', C2] token-major."""
    _lib.require_device()
    _check(x, torch.float32, "x")
    B, F, T = x.shape
    C2, KP = Wc.shape
    Fo, To = (F - kc) // stride + 1, (T - kc) // stride + 1
    out = torch.empty((B, To, Fo, C2), device=x.device, dtype=out_dtype)
    call("b2c_convsub_fwd", ptr(x), ptr(Wc), ptr(bc), ptr(out), int(out_dtype == torch.float32), B, F, T, C2, kc, stride,
         KP, stream_ptr())
    return out


def convsub_wgrad(x, dout, dWc, kc: int, stride: int):
    B, F, T = x.shape
    C2, KP = dWc.shape
    call("b2c_convsub_wgrad", ptr(x), ptr(dout), ptr(dWc), B, F, T, C2, kc, stride, KP, stream_ptr())
    return dWc


def convsub_compose_bwd(dWc, dbc, W1, b1, W2, s1: int, dW1, db1, dW2, db2):
    C1, _, k1, _ = W1.shape
    C2, _, k2, _ = W2.shape
    call("b2c_convsub_compose_bwd", ptr(dWc), ptr(dbc), ptr(W1), ptr(b1), ptr(W2), C1, C2, k1, s1, k2, dWc.shape[1],
         ptr(dW1), ptr(db1), ptr(dW2), ptr(db2), stream_ptr())


def specaug_apply(x, tau_i32, warp_w, warp_w0, fmask, tmask, apply_masks: bool, mask_value: float):
    """x fp32 [B,F,T]; parameter tensors are int32 device tensors (or None when the count is 0)."""
    _lib.require_device()
    _check(x, torch.float32, "x")
    B, F, T = x.shape
    out = torch.empty_like(x)
    n_warp = 0 if warp_w is None else warp_w.shape[0]
    n_f = 0 if fmask is None else fmask.shape[0]
    n_t = 0 if tmask is None else tmask.shape[0]
    call("b2c_specaug_apply", ptr(x), ptr(out), ptr(tau_i32), ptr(warp_w), ptr(warp_w0), n_warp, ptr(fmask), n_f,
         ptr(tmask), n_t, int(apply_masks), float(mask_value), B, F, T, stream_ptr())
    return out


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None):
    _lib.require_device()
    _check(src, torch.float32, "src")
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    call("b2c_cast_f32_bf16", ptr(src), ptr(dst), src.numel(), stream_ptr())
    return dst


def scale_cast(src, alpha: float, drop_p: float = 0.0, drop_seed: int = 0, colsum: Optional[torch.Tensor] = None):
    rows, cols = src.shape
    dst = torch.empty((rows, cols), device=src.device, dtype=torch.bfloat16)
    call("b2c_scale_cast", ptr(src), ptr(dst), rows, cols, alpha, drop_p, drop_seed, ptr(colsum), stream_ptr())
    return dst


def colsum(a, out):
    rows, cols = a.shape
    call("b2c_colsum_bf16", ptr(a), a.stride(0), rows, cols, ptr(out), stream_ptr())
    return out


def repack_heads(src, nblk: int, src_dh: int, dst_dh: int):
    rows = src.shape[0]
    dst = torch.empty((rows, nblk * dst_dh), device=src.device, dtype=torch.bfloat16)
    call("b2c_repack_heads", ptr(src), ptr(dst), rows, nblk, src_dh, dst_dh, stream_ptr())
    return dst


def convmod_bn_bwd_reduce(dc, y, mean, rstd, bn_w, bn_b, sums=None):
    rows, d = y.shape
    if sums is None:
        sums = torch.zeros(2 * d, device=y.device, dtype=torch.float64)
    call("b2c_convmod_bn_bwd_reduce", ptr(dc), ptr(y), ptr(mean), ptr(rstd), ptr(bn_w), ptr(bn_b), ptr(sums), rows, d,
         stream_ptr())
    return sums


def convmod_dw_bwd(dc, y, u, mean, rstd, bn_w, bn_b, sums, bn_train: bool, wdw, dwdw, dbdw, dbn_w, dbn_b, B: int, T: int,
                   dbias_u=None):
    """dbias_u (fp32 [2d], optional): accumulates the column sums of du (bias gradient of the pointwise conv before the GLU)."""
    d, k = wdw.shape[0], wdw.shape[-1]
    du = torch.empty((B * T, 2 * d), device=u.device, dtype=torch.bfloat16)
    call("b2c_convmod_dw_bwd", ptr(dc), ptr(y), ptr(u), ptr(mean), ptr(rstd), ptr(bn_w), ptr(bn_b), ptr(sums),
         int(bn_train), ptr(wdw), ptr(du), ptr(dwdw), ptr(dbdw), ptr(dbn_w), ptr(dbn_b), ptr(dbias_u), B, T, d, k,
         stream_ptr())
    return du


def attention_bwd(qkv, lengths_i32, att, datt, lse, B: int, T: int, H: int, scale: float, drop_p: float = 0.0,
                  drop_seed: int = 0):
    dqkv = torch.empty_like(qkv)
    delta = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32)
    dq32 = torch.empty((B * T, H * 64), device=qkv.device, dtype=torch.float32)   # zeroed by the library, accumulated by TMA reduce-adds
    call("b2c_attention_bwd", ptr(qkv), qkv.stride(0), H * 64, 2 * H * 64, ptr(lengths_i32), ptr(att), ptr(datt),
         att.stride(0), ptr(lse), ptr(delta), ptr(dq32), ptr(dqkv), B, T, H, scale, drop_p, drop_seed, stream_ptr(),
         flops=14.0 * B * H * T * T * 64)
    return dqkv
