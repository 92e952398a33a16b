"""Tiny driver for ncu: conv-module chain and attention kernels at Conformer-M shapes."""
import sys
import torch
sys.path.insert(0, ".")
from nn_conformer_for_speech_recognition_b200 import ops
M, d, H, T, B, k = 23872, 256, 4, 373, 64, 31
bf = torch.bfloat16
dev = "cuda"
u = torch.randn(M, 2 * d, device=dev).to(bf)
wdw, bdw = torch.randn(d, k, device=dev) / 5, torch.zeros(d, device=dev)
one, zero = torch.ones(d, device=dev), torch.zeros(d, device=dev)
qkv = torch.randn(M, 3 * H * 64, device=dev).to(bf)
lens = torch.full((B,), T, device=dev, dtype=torch.int32)
dc = torch.randn(M, d, device=dev).to(bf)
for _ in range(2):
    ops.convmod_eval(u, wdw, bdw, one, zero, zero, one, B, T)
    yy, stats = ops.convmod_dw_fwd(u, wdw, bdw, B, T)
    mean_c, rstd_c = ops.convmod_bn_finalize(stats, M, None, None)
    ops.convmod_bn_apply(yy, mean_c, rstd_c, one, zero)
    sums = ops.convmod_bn_bwd_reduce(dc, yy, mean_c, rstd_c, one, zero)
    dwdw, dbdw, dg, db = torch.zeros(d, k, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    ops.convmod_dw_bwd(dc, yy, u, mean_c, rstd_c, one, zero, sums, True, wdw, dwdw, dbdw, dg, db, B, T)
    att, lse = ops.attention_fwd(qkv, lens, B, T, H, 0.125, save_lse=True, drop_p=0.5, drop_seed=1)
    ops.attention_bwd(qkv, lens, att, torch.randn_like(att), lse, B, T, H, 0.125, 0.5, 1)
torch.cuda.synchronize()
print("ok")
