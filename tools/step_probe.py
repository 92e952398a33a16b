"""Driver for `ncu --set full`: one launch of every distinct hot kernel of the training step at Conformer-M shapes
(64 x 15 s): forward/backward GEMM variants with their training epilogues, attention, LayerNorm, conv module."""
import sys
import torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
M, d, f, H, T, B, k = 23872, 256, 1024, 4, 373, 64, 31
bf, dev = torch.bfloat16, "cuda"
a = torch.randn(M, d, device=dev).to(bf); hbuf = torch.randn(M, f, device=dev).to(bf)
w1 = torch.randn(f, d, device=dev).to(bf); w2 = torch.randn(d, f, device=dev).to(bf); wq = torch.randn(3 * d, d, device=dev).to(bf)
b1, b2 = torch.randn(f, device=dev), torch.randn(d, device=dev)
x = torch.randn(M, d, device=dev); u2 = torch.empty(M, f, device=dev, dtype=bf); out_f = torch.empty(M, f, device=dev, dtype=bf)
cs = torch.zeros(f, device=dev); dW = torch.zeros(d, f, device=dev)
g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
qkv = torch.randn(M, 3 * H * 64, device=dev).to(bf); lens = torch.full((B,), T, device=dev, dtype=torch.int32)
u = torch.randn(M, 2 * d, device=dev).to(bf); wdw, bdw = torch.randn(d, k, device=dev) / 5, torch.zeros(d, device=dev)
one, zero = torch.ones(d, device=dev), torch.zeros(d, device=dev)
for _ in range(2):
    ops.gemm(a, w1, M=M, N=f, K=d, bias=b1, act=ops.ACT_SILU, out=out_f, out2=u2, drop_p=0.5, drop_seed=1)     # FFN-1 fwd (train)
    ops.gemm(hbuf, w2, M=M, N=d, K=f, bias=b2, resid=x, out_dtype=torch.float32, alpha=0.5, drop_p=0.5, drop_seed=2)  # FFN-2 fwd
    ops.gemm(a, wq, M=M, N=3 * d, K=d, bias=None)                                                             # QKV
    ops.gemm(a, w2, M=M, N=f, K=d, b_mn=True, act=ops.ACT_SILU_BWD, aux=u2, out=out_f, alpha=2.0, colsum=cs)      # dgrad SiLU' (mask from u2)
    ops.gemm(hbuf, w1, M=M, N=d, K=f, b_mn=True)                                                             # dgrad K=1024
    ops.gemm(a, hbuf, M=d, N=f, K=M, a_mn=True, b_mn=True, out=dW, accumulate=True, split_k=37)              # wgrad
    y, mean, rstd = ops.layernorm_fwd(x, g, be, save_stats=True)
    dx, dg, db, cs2 = torch.zeros_like(x), torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.layernorm_bwd(y, x, g, mean, rstd, dx, True, dg, db, cast=(0.5, 0.5, 3, cs2))
    att, lse = ops.attention_fwd(qkv, lens, B, T, H, 0.125, save_lse=True, drop_p=0.5, drop_seed=1)
    ops.attention_bwd(qkv, lens, att, torch.randn_like(att), lse, B, T, H, 0.125, 0.5, 1)
    ops.convmod_eval(u, wdw, bdw, one, zero, zero, one, B, T)
    yy, stats = ops.convmod_dw_fwd(u, wdw, bdw, B, T)
    mc, rc = ops.convmod_bn_finalize(stats, M, None, None)
    ops.convmod_bn_apply(yy, mc, rc, one, zero)
    sums = ops.convmod_bn_bwd_reduce(y, yy, mc, rc, one, zero)
    ops.convmod_dw_bwd(y, yy, u, mc, rc, one, zero, sums, True, wdw, torch.zeros(d, k, device=dev), torch.zeros(d, device=dev), dg, db, B, T)
torch.cuda.synchronize()
print("ok")
