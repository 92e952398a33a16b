"""One eager launch (after one warm-up launch) of every hot kernel at the Conformer-M benchmark shapes, with the tile shapes the
training step uses -- the target of the `ncu --set full` capture:
    gpurun -- 'python tools/ncu_targets.py > gpurun_out/t.log 2>&1 && ncu --set full --clock-control none --import-source on
               --kernel-name-base demangled -k regex:b2c:: -o gpurun_out/r02_full python tools/ncu_targets.py'"""
import math, sys, torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
from nn_conformer_for_speech_recognition_b200.ctc import ctc_loss_from_logits
dev, bf = "cuda", torch.bfloat16
M, d, f, H, T, B, k = 23872, 256, 1024, 4, 373, 64, 31
a_d = torch.randn(M, d, device=dev).to(bf); a_f = torch.randn(M, f, device=dev).to(bf)
w_fd = (torch.randn(f, d, device=dev) / 16).to(bf); w_df = (torch.randn(d, f, device=dev) / 32).to(bf)
w_dd = (torch.randn(d, d, device=dev) / 16).to(bf); w_3d = (torch.randn(3 * d, d, device=dev) / 16).to(bf)
bias_f, bias_d = torch.randn(f, device=dev), torch.randn(d, device=dev)
x = torch.randn(M, d, device=dev)
u_f = torch.empty(M, f, device=dev, dtype=bf)
g, be = torch.ones(d, device=dev), torch.zeros(d, device=dev)
qkv = torch.randn(M, 3 * H * 64, device=dev).to(bf)
lens = torch.full((B,), T, device=dev, dtype=torch.int32)
uu = torch.randn(M, 2 * d, device=dev).to(bf)
wdw, bdw = torch.randn(d, k, device=dev) / 5, torch.zeros(d, device=dev)
one, zero = torch.ones(d, device=dev), torch.zeros(d, device=dev)
dW = torch.zeros(d, f, device=dev); dWs = torch.zeros(d, d, device=dev)
for rep in range(2):
    h = ops.gemm(a_d, w_fd, M=M, N=f, K=d, bias=bias_f, act=ops.ACT_SILU, out2=u_f, drop_p=0.5, drop_seed=3)          # FFN1 fwd (train)
    ops.gemm(a_f, w_df, M=M, N=d, K=f, bias=bias_d, resid=x, out_dtype=torch.float32, alpha=0.5, drop_p=0.5, drop_seed=4)  # FFN2 fwd
    ops.gemm(a_d, w_3d, M=M, N=3 * d, K=d, bias=None)                                                                # QKV
    ops.gemm(a_d, w_dd, M=M, N=d, K=d, bias=bias_d, resid=x, out_dtype=torch.float32, drop_p=0.5, drop_seed=5)          # out-proj / pw2
    ops.gemm(a_d, w_df, M=M, N=f, K=d, b_mn=True, act=ops.ACT_SILU_BWD, aux=u_f, alpha=2.0)                           # SiLU' dgrad
    ops.gemm(a_f, w_fd, M=M, N=d, K=f, b_mn=True)                                                                   # dgrad K=1024
    ops.gemm(a_d, a_f, M=d, N=f, K=M, a_mn=True, b_mn=True, out=dW, accumulate=True, split_k=18)                    # wgrad 256x1024
    ops.gemm(a_d, a_d, M=d, N=d, K=M, a_mn=True, b_mn=True, out=dWs, accumulate=True, split_k=74)                   # wgrad 256x256
    y, mean, rstd = ops.layernorm_fwd(x, g, be, save_stats=True)
    dx, dg, db = torch.zeros_like(x), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.layernorm_bwd(y, x, g, mean, rstd, dx, True, dg, db, cast=(0.5, 0.5, 7, torch.zeros(d, device=dev)))
    att, lse = ops.attention_fwd(qkv, lens, B, T, H, 0.125, save_lse=True, drop_p=0.5, drop_seed=1)
    ops.attention_bwd(qkv, lens, att, torch.randn_like(att), lse, B, T, H, 0.125, 0.5, 1)
    yy, stats = ops.convmod_dw_fwd(uu, wdw, bdw, B, T)
    mean_c, rstd_c = ops.convmod_bn_finalize(stats, M, None, None)
    c = ops.convmod_bn_apply(yy, mean_c, rstd_c, one, zero)
    sums = ops.convmod_bn_bwd_reduce(c, yy, mean_c, rstd_c, one, zero)
    ops.convmod_dw_bwd(c, yy, uu, mean_c, rstd_c, one, zero, sums, True, wdw, torch.zeros(d, k, device=dev), torch.zeros(d, device=dev), dg, db, B, T)
    c2, mean_c2, rstd_c2 = ops.convmod_bn_finalize_apply(yy, stats, M, one, zero, None, None)                         # finalize + apply, one launch
    y1, m1, r1, a2, m2, r2 = ops.layernorm_fwd_pair(x, g, be, g, be, save_stats=True)                                   # LayerNorm pair fwd
    dxp = torch.empty_like(x)
    ops.layernorm_bwd_pair(a2, y1, g, m2, r2, dg, db, dx, x, g, m1, r1, dg, db, dxp, cast=(0.5, 0.5, 9, torch.zeros(d, device=dev)))   # pair bwd
    W1c, b1c = torch.randn(512, 1, 7, 7, device=dev) / 7, torch.zeros(512, device=dev)
    W2c, b2c_ = torch.randn(128, 512, 3, 3, device=dev) / 68, torch.zeros(128, device=dev)
    Wc, bc = ops.convsub_compose(W1c, b1c, W2c, b2c_, 2, 128)                                                          # staged composition
    ops.convsub_compose_bwd(torch.randn(128, 128, device=dev), torch.randn(128, device=dev), W1c, b1c, W2c, 2,
                            torch.zeros_like(W1c), torch.zeros_like(b1c), torch.zeros_like(W2c), torch.zeros_like(b2c_))
    logits = torch.randn(B, T, 1024, device=dev, requires_grad=True)
    ctc_loss_from_logits(logits, torch.randint(1, 1024, (B, 20), device=dev), [T] * B, [20] * B).backward()
torch.cuda.synchronize()
print("ok")
