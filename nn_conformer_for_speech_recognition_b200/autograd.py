"""Autograd glue for the Conformer stack: ONE torch.autograd.Function spans all layers, so forward and backward
are straight sequences of kernel launches (capturable in a CUDA graph) and torch's autograd engine sees a single
node.  Gradient math mirrors what autograd derives for torchaudio's modules (conformer.py:18-212); every
matrix product is the tcgen05 GEMM in one of its three operand-major modes:

    forward   Y  = X W^T            (K-major, K-major)
    dgrad     dX = dY W             (K-major, MN-major)  -- W used as stored, no transpose copy
    wgrad     dW = dY^T X           (MN-major, MN-major) -- split over tokens, fp32 atomics into the grad arena
"""
from __future__ import annotations

import math
import os
from typing import List

import torch

from . import _lib, ops
from .conformer import _LN_PAIR, HEAD_PAD, Conformer, _get, run_layers


def conformer_apply(model: Conformer, inp: torch.Tensor, lengths_i32: torch.Tensor) -> torch.Tensor:
    B, T, d = inp.shape
    params = [_get(layer, n) for layer in model.conformer_layers for n in model.layer_params]
    need_grad = torch.is_grad_enabled() and (inp.requires_grad or any(p.requires_grad for p in params))
    drop_p = model.dropout_p if model.training else 0.0
    if not need_grad:
        x = inp.detach().reshape(B * T, d).to(torch.float32).clone(memory_format=torch.contiguous_format)
        if drop_p > 0.0:
            # train mode without autograd (torch.no_grad(): MC-dropout, noisy-student inference): torchaudio still applies
            # dropout there, so run the dropout-enabled pipeline and drop what it saves for a backward pass
            with _lib.epoch_scope(_lib.epoch_snapshot(inp.device)):
                y = run_layers(model, x, lengths_i32, B, T, [], True, drop_p, model.next_seed())
        else:
            y = run_layers(model, x, lengths_i32, B, T, None, model.training, 0.0, model.next_seed())
        return y.view(B, T, d)
    return _ConformerFn.apply(inp, lengths_i32, model, *params)


# Weight-gradient GEMMs have no consumer inside the backward pass: they run on a side stream (a parallel branch when the
# backward pass is captured into a CUDA graph), where their CTAs and memory traffic fill what the dependent chain (dgrad GEMM ->
# LayerNorm backward -> cast ...) leaves idle.  Measured on the final round-2 build: 15.07 -> 14.53 ms per training step
# (round 1 had measured it neutral; with programmatic dependent launch the two branches now interleave).  B2C_WGRAD_STREAM=0 disables.
_WGRAD_SIDE_STREAM = os.environ.get("B2C_WGRAD_STREAM", "1") != "0"
_SIDE = {}


def _side_stream(dev) -> torch.cuda.Stream:
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


def _split_k(tokens: int, out_tiles: int, sms: int = 148) -> int:
    """split factor for a weight-gradient GEMM: ONE wave of CTAs (`out_tiles` = 128 x 256 output tiles), at least 4
    K-blocks each (measured, tools/wgrad_probe.py: more splits only add fp32 reductions)."""
    kb = (tokens + 63) // 64
    want = max(1, sms // max(1, out_tiles))
    return max(1, min(want, kb // 4 if kb >= 4 else 1))


def defer_first_layernorm(li: int, d: int, exchange_active: bool, group_lo) -> bool:
    """Backward-pass scheduling rule of the LayerNorm pairs: layer `li` may leave the backward of its first LayerNorm
    (ffn1.sequential.0) to the pair kernel that opens layer li-1's backward pass -- unless there is no layer below, the
    width has no pair instantiation worth running (d > 512: 255 registers, spills), or `li` is the first layer of a
    gradient-exchange group: that group's span of the gradient arena is handed to the all-reduce at the END of layer li's
    backward pass and must be final by then (dp.GradSync.attach; `group_lo` maps a group's first layer to one past its last)."""
    if not _LN_PAIR or li <= 0 or d > 512:
        return False
    return not (exchange_active and li in group_lo)


class _ConformerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, lengths_i32, model, *params):
        B, T, d = inp.shape
        x = inp.detach().reshape(B * T, d).to(torch.float32).clone(memory_format=torch.contiguous_format)
        state: List[dict] = []
        drop_p = model.dropout_p if model.training else 0.0
        seed = model.next_seed()
        # one epoch word per CALL (device step counter + 1, copied by a capturable launch): forward and backward of this
        # call read the same word even when other training forwards run in between (advisor finding, round 1)
        ctx.epoch = _lib.epoch_snapshot(inp.device) if drop_p > 0.0 else None
        # gradients are wanted: always run the two-pass conv-module chain so `y` exists for the backward pass
        with _lib.epoch_scope(ctx.epoch):
            y = run_layers(model, x, lengths_i32, B, T, state, model.training, drop_p, seed, force_two_pass=True)
        ctx.model, ctx.state, ctx.lengths = model, state, lengths_i32
        if getattr(model, "_debug_keep_state", False):          # tests: saved activations / seed / epoch of this call
            model._debug_state = dict(state=list(state), seed=seed, epoch=ctx.epoch, drop_p=drop_p)
        ctx.weights = model._cache.bf16          # the operands this forward used (static buffers; no re-cast in backward)
        ctx.dims = (B, T, d)
        ctx.drop_p, ctx.seed = drop_p, seed
        ctx.in_dtype = inp.dtype
        return y.view(B, T, d)

    @staticmethod
    def backward(ctx, dy):
        with _lib.epoch_scope(ctx.epoch), _lib.nvtx("conformer.backward"):
            return _ConformerFn._backward(ctx, dy)

    @staticmethod
    def _backward(ctx, dy):
        model: Conformer = ctx.model
        B, T, d = ctx.dims
        M = B * T
        f, H, k = model.ffn_dim, model.num_heads, model.kernel_size
        Hd = H * HEAD_PAD
        dh = model.head_dim
        scale = 1.0 / math.sqrt(dh)
        drop_p, seed = ctx.drop_p, ctx.seed
        keep_scale = float(_lib.load().b2c_dropout_keep_scale(drop_p)) if drop_p > 0 else 1.0
        weights = ctx.weights
        dev = dy.device
        dx = dy.reshape(M, d).to(torch.float32).contiguous()
        if dx.data_ptr() == dy.data_ptr():
            dx = dx.clone()

        # one zeroed fp32 arena holds every parameter gradient of the stack (atomics accumulate into it)
        rel = model.rel_att
        nq = 4 if rel else 3                  # rel_att: the in-projection carries a 4th slab (q + pos_bias_v)
        padded = dh != HEAD_PAD or rel
        _LAYER_PARAMS = model.layer_params
        sizes = []
        for layer in model.conformer_layers:
            for n in _LAYER_PARAMS:
                p = _get(layer, n)
                if padded and n == "self_attn.in_proj_weight":
                    sizes.append(nq * Hd * d)
                elif padded and n == "self_attn.in_proj_bias":
                    sizes.append(nq * Hd)
                elif padded and n in ("self_attn.out_proj.weight", "self_attn.linear_pos.weight"):
                    sizes.append(d * Hd)
                elif n in ("self_attn.pos_bias_u", "self_attn.pos_bias_v"):
                    sizes.append(4)           # filled from the in-projection bias gradient below
                else:
                    sizes.append(p.numel())
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + ((s + 3) // 4) * 4)         # keep every slice 16-byte aligned
        arena = torch.zeros(offs[-1], device=dev, dtype=torch.float32)
        sums_all = torch.zeros((model.num_layers, 2 * d), device=dev, dtype=torch.float64)   # BatchNorm backward sums, one fill
        np_ = len(_LAYER_PARAMS)

        def G(li, name):
            i = li * np_ + _LAYER_PARAMS.index(name)
            return arena[offs[i]:offs[i] + sizes[i]]

        # Weight-gradient GEMMs have no consumer inside the backward pass: with B2C_WGRAD_STREAM=1 they go to a side stream,
        # where their CTAs fill the SMs the dependent chain (dgrad GEMM -> LayerNorm backward -> cast ...) leaves idle in
        # its tail waves (+2 % when the GEMMs still had ragged last waves; neutral since, hence off by default).
        # Operands stay referenced until the streams are joined again.
        main = torch.cuda.current_stream()
        side = _side_stream(dev) if _WGRAD_SIDE_STREAM else None
        # data-parallel training: dp.GradSync.attach() asks for the arena to be all-reduced in `groups` spans of layers
        gsync = getattr(model, "_grad_sync", None)
        group_lo = {}
        if gsync is not None and gsync.world > 1:
            ng, Lr = model._grad_sync_groups, model.num_layers
            bounds = [round(i * Lr / ng) for i in range(ng + 1)]
            group_lo = {bounds[i]: bounds[i + 1] for i in range(ng)}      # first layer of a group -> one past its last
        else:
            gsync = None
        keep = []

        def wgrad(dyb, act, Mo, No, out):
            """out[Mo,No] += dyb[tokens,Mo]^T act[tokens,No]"""
            tiles = ((Mo + 127) // 128) * ((No + 255) // 256)
            if side is None:
                ops.gemm(dyb, act, M=Mo, N=No, K=M, a_mn=True, b_mn=True, out=out.view(Mo, No), accumulate=True,
                         split_k=_split_k(M, tiles))
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.gemm(dyb, act, M=Mo, N=No, K=M, a_mn=True, b_mn=True, out=out.view(Mo, No), accumulate=True,
                         split_k=_split_k(M, tiles))
            keep.append((dyb, act))

        pending = None
        for li in reversed(range(model.num_layers)):
            layer = model.conformer_layers[li]
            st = ctx.state[li]
            w = weights[li]
            sd = (seed + li * 16) if drop_p > 0 else 0
            P = lambda name: _get(layer, name)                                   # noqa: E731

            # ---- 5. final LayerNorm: the stream passes through it (no skip connection)
            # every LayerNorm backward below also emits the NEXT sub-layer's starting operand: bf16(alpha * dropout(dx)) of
            # the residual branch it is about to enter, with that branch's output-bias gradient (column sums) -- one
            # pass over dx instead of LayerNorm-backward followed by a separate scale/cast kernel
            s5 = st["final"]
            dx_new = torch.empty_like(dx)
            if pending is not None:
                # the layer above left its first LayerNorm's backward undone: both LayerNorms in one pass over the rows
                # (b2c_layernorm_bwd_pair: the intermediate gradient is neither written nor re-read)
                da2, y1, g2, m2, r2, dg2, db2 = pending
                pending = None
                dyb = ops.layernorm_bwd_pair(da2, y1, g2, m2, r2, dg2, db2, dx, s5["x"], P("final_layer_norm.weight"),
                                             s5["mean"], s5["rstd"], G(li, "final_layer_norm.weight"),
                                             G(li, "final_layer_norm.bias"), dx_new,
                                             cast=(0.5, drop_p, sd + 12 + 1, G(li, "ffn2.sequential.4.bias")))
            else:
                dyb = ops.layernorm_bwd(dx, s5["x"], P("final_layer_norm.weight"), s5["mean"], s5["rstd"], dx_new, False,
                                        G(li, "final_layer_norm.weight"), G(li, "final_layer_norm.bias"),
                                        cast=(0.5, drop_p, sd + 12 + 1, G(li, "ffn2.sequential.4.bias")))
            dx = dx_new

            def ffn_bwd(pre, site, dyb, next_cast, defer=False):
                s = st[pre]
                wgrad(dyb, s["h"], d, f, G(li, pre + ".sequential.4.weight"))
                # s["u"] holds B2C_DROPPED_PREACT where the forward dropped the activation (SiLU' = 0 there), so the mask is
                # not regenerated: kept elements only need the 1/keep scale
                du = ops.gemm(dyb, w[pre + ".sequential.4.weight"], M=M, N=f, K=d, b_mn=True, act=ops.ACT_SILU_BWD,
                              aux=s["u"], alpha=keep_scale, colsum=G(li, pre + ".sequential.1.bias"))
                wgrad(du, s["a"], f, d, G(li, pre + ".sequential.1.weight"))
                da = ops.gemm(du, w[pre + ".sequential.1.weight"], M=M, N=d, K=f, b_mn=True)
                if defer:          # this LayerNorm's backward runs fused with the final LayerNorm of the layer below
                    return (da, s["x"], P(pre + ".sequential.0.weight"), s["mean"], s["rstd"],
                            G(li, pre + ".sequential.0.weight"), G(li, pre + ".sequential.0.bias"))
                return ops.layernorm_bwd(da, s["x"], P(pre + ".sequential.0.weight"), s["mean"], s["rstd"], dx, True,
                                         G(li, pre + ".sequential.0.weight"), G(li, pre + ".sequential.0.bias"),
                                         cast=next_cast)

            # ---- 4. FFN2 (leaves the conv module's starting operand behind)
            dyb = ffn_bwd("ffn2", 12, dyb, (1.0, drop_p, sd + 8, G(li, "conv_module.sequential.5.bias")))

            # ---- 3. conv module
            s = st["conv"]
            bn = layer.conv_module.sequential[3]
            wgrad(dyb, s["c"], d, d, G(li, "conv_module.sequential.5.weight"))
            dc = ops.gemm(dyb, w["conv_module.sequential.5.weight"], M=M, N=d, K=d, b_mn=True)
            sums = ops.convmod_bn_bwd_reduce(dc, s["y"], s["bmean"], s["brstd"], bn.weight, bn.bias, sums=sums_all[li])
            du = ops.convmod_dw_bwd(dc, s["y"], s["u"], s["bmean"], s["brstd"], bn.weight, bn.bias, sums, s["bn_train"],
                                    P("conv_module.sequential.2.weight").view(d, k),
                                    G(li, "conv_module.sequential.2.weight"), G(li, "conv_module.sequential.2.bias"),
                                    G(li, "conv_module.sequential.3.weight"), G(li, "conv_module.sequential.3.bias"), B, T,
                                    dbias_u=G(li, "conv_module.sequential.0.bias"))
            wgrad(du, s["a"], 2 * d, d, G(li, "conv_module.sequential.0.weight"))
            da = ops.gemm(du, w["conv_module.sequential.0.weight"], M=M, N=d, K=2 * d, b_mn=True)
            dyb = ops.layernorm_bwd(da, s["x"], P("conv_module.layer_norm.weight"), s["mean"], s["rstd"], dx, True,
                                    G(li, "conv_module.layer_norm.weight"), G(li, "conv_module.layer_norm.bias"),
                                    cast=(1.0, drop_p, sd + 5, G(li, "self_attn.out_proj.bias")))

            # ---- 2. attention
            s = st["attn"]
            wgrad(dyb, s["att"], d, Hd, G(li, "self_attn.out_proj.weight"))
            datt = ops.gemm(dyb, w["self_attn.out_proj.weight"], M=M, N=Hd, K=d, b_mn=True)
            if rel:
                dqkv, dpw = ops.attention_rel_bwd(s["qkv"], s["pw"], ctx.lengths, s["att"], datt, s["lse"], B, T, H, scale,
                                                  drop_p, sd + 4)
                pe = model.rel_table(T, dev)
                # Pw = PE W_pos^T  =>  dW_pos[Hd, d] = dPw^T PE
                ops.gemm(ops.cast_bf16(dpw), pe, M=Hd, N=d, K=pe.shape[0], a_mn=True, b_mn=True,
                         out=G(li, "self_attn.linear_pos.weight").view(Hd, d), accumulate=True)
            else:
                dqkv = ops.attention_bwd(s["qkv"], ctx.lengths, s["att"], datt, s["lse"], B, T, H, scale, drop_p, sd + 4)
            if side is None:
                ops.colsum(dqkv, G(li, "self_attn.in_proj_bias"))
            else:                                   # the in-projection bias gradient has no consumer in the backward pass either
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.colsum(dqkv, G(li, "self_attn.in_proj_bias"))
                keep.append((dqkv,))
            wgrad(dqkv, s["a"], nq * Hd, d, G(li, "self_attn.in_proj_weight"))
            da = ops.gemm(dqkv, w["self_attn.in_proj_weight"], M=M, N=d, K=nq * Hd, b_mn=True)
            dyb = ops.layernorm_bwd(da, s["x"], P("self_attn_layer_norm.weight"), s["mean"], s["rstd"], dx, True,
                                    G(li, "self_attn_layer_norm.weight"), G(li, "self_attn_layer_norm.bias"),
                                    cast=(0.5, drop_p, sd + 0 + 1, G(li, "ffn1.sequential.4.bias")))

            # ---- 1. FFN1 (its LayerNorm is deferred into the pair kernel of the layer below, unless this layer's gradients
            # are about to be handed to the gradient exchange: a span must be final when it leaves)
            defer = defer_first_layernorm(li, d, gsync is not None, group_lo)
            r = ffn_bwd("ffn1", 0, dyb, None, defer)
            pending = r if defer else None
            ctx.state[li] = None                  # free this layer's saved activations
            if gsync is not None and li in group_lo:
                # every gradient of layers [li, group end) is final: hand the arena span to the gradient exchange now, so
                # the all-reduce runs (NCCL stream / parallel graph branch) under the backward kernels of the layers below
                if side is not None:
                    main.wait_stream(side)
                hi_layer = group_lo[li]
                gsync.reduce_span(arena[offs[li * np_]:offs[hi_layer * np_]])

        if side is not None:
            main.wait_stream(side)
            keep.clear()
        if gsync is not None:
            gsync.join()                          # the reduced spans are visible to the stream before the gradients are handed out
        grads = []
        for li, layer in enumerate(model.conformer_layers):
            for n in _LAYER_PARAMS:
                p = _get(layer, n)
                g = G(li, n)
                if padded and n == "self_attn.in_proj_weight":
                    g4 = g.view(nq, H, HEAD_PAD, d)[:, :, :dh]
                    g = g4[:3].clone() if rel else g4
                    if rel:
                        g[0] += g4[3]             # W_q feeds both q+u and q+v
                    g = g.reshape(3 * d, d)
                elif padded and n == "self_attn.in_proj_bias":
                    g4 = g.view(nq, H, HEAD_PAD)[:, :, :dh]
                    g = g4[:3].clone() if rel else g4
                    if rel:
                        g[0] += g4[3]
                    g = g.reshape(3 * d)
                elif n == "self_attn.pos_bias_u":      # d(q+u) summed over tokens = slab 0 of the in-projection bias gradient
                    g = G(li, "self_attn.in_proj_bias").view(nq, H, HEAD_PAD)[0, :, :dh].clone()
                elif n == "self_attn.pos_bias_v":
                    g = G(li, "self_attn.in_proj_bias").view(nq, H, HEAD_PAD)[3, :, :dh].clone()
                elif padded and n == "self_attn.linear_pos.weight":
                    g = g.view(H, HEAD_PAD, d)[:, :dh].reshape(d, d)
                elif padded and n == "self_attn.out_proj.weight":
                    g = g.view(d, H, HEAD_PAD)[:, :, :dh].reshape(d, d)
                else:
                    g = g.view(p.shape)
                grads.append(g if p.requires_grad else None)
        dinp = dx.view(B, T, d).to(ctx.in_dtype) if ctx.needs_input_grad[0] else None
        return (dinp, None, None, *grads)
